// Envelope multiply, batched determinants (LU / Gauss-Jordan with partial pivoting, inverse-trace
// Laplacian contraction) and the signed log-sum-exp that ends in log psi^2 and E_loc.
// Reference: model/orbitals/envelope_orbitals.py:39-127, model/wavefunction.py:63-83,
// hamiltonian.py:206-216 (forward-Laplacian kinetic energy), :34-39 (potential).
#include <cstdlib>
#include "dpe_internal.cuh"

namespace dpe {

// ------------------------------------------------------------------------------------------------
// mo[b,i,c,col] = env[b,i,col] (x) bf[b,i,c,col]  in place (product rule; env depends on r_i only)
// env = sum_J w[J,col] exp(-softplus(alpha[J,col]) |r_i - R_J|)      col = det*N + orb
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_envelope(const float *__restrict__ r, const float *__restrict__ R, int Bc, int N,
                                                   int U, int I, int C, int cols, const float *__restrict__ spa_up,
                                                   const float *__restrict__ spa_dn, const float *__restrict__ w_up,
                                                   const float *__restrict__ w_dn, float *__restrict__ mo) {
    const long total = (long)Bc * N * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int col = idx % cols;
        long bi = idx / cols;
        int i = bi % N;
        const float *spa = i < U ? spa_up : spa_dn;
        const float *wt = i < U ? w_up : w_dn;
        const float *ri = r + bi * 3;
        float env = 0.f, e1[3] = {0.f, 0.f, 0.f}, el = 0.f;
        for (int J = 0; J < I; ++J) {
            float dx = ri[0] - R[J * 3], dy = ri[1] - R[J * 3 + 1], dz = ri[2] - R[J * 3 + 2];
            float d = sqrtf(dx * dx + dy * dy + dz * dz);
            float a = spa[(long)J * cols + col];
            float e = __fmul_rn(wt[(long)J * cols + col], expf(-a * d));   // no FMA contraction: same bits in both modes
            env = __fadd_rn(env, e);
            if (C > 1) {
                float inv = 1.f / d;
                float g = -a * e * inv;
                e1[0] = fmaf(g, dx, e1[0]); e1[1] = fmaf(g, dy, e1[1]); e1[2] = fmaf(g, dz, e1[2]);
                el = fmaf(e, a * a - 2.f * a * inv, el);
            }
        }
        float *p = mo + bi * (long)C * cols + col;
        if (C == 1) {
            p[0] *= env;
        } else {
            const float bf0 = p[0];
            const int ci = 1 + 3 * i;
            const float t0 = p[(long)ci * cols], t1 = p[(long)(ci + 1) * cols], t2 = p[(long)(ci + 2) * cols];
            const float lap_extra = el * bf0 + 2.f * (e1[0] * t0 + e1[1] * t1 + e1[2] * t2);
            constexpr int UB = 8;      // batch the loads ahead of the in-place stores
            for (int c0 = 0; c0 < C; c0 += UB) {
                float v[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (c0 + u < C) v[u] = p[(long)(c0 + u) * cols];
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int c = c0 + u;
                    if (c < C) {
                        float o = v[u] * env;
                        if (c >= ci && c < ci + 3) o += (c == ci ? e1[0] : (c == ci + 1 ? e1[1] : e1[2])) * bf0;
                        if (c == C - 1) o += lap_extra;
                        p[(long)c * cols] = o;
                    }
                }
            }
        }
    }
}

// Forward pass (one channel): one block per (walker, electron) row, the el-ion distances once per block, no index divisions.
// Same arithmetic and summation order as k_envelope.
__global__ void __launch_bounds__(128) k_envelope_fwd(const float *__restrict__ r, const float *__restrict__ R, int N, int U, int I, int cols,
                                                       const float *__restrict__ spa_up, const float *__restrict__ spa_dn,
                                                       const float *__restrict__ w_up, const float *__restrict__ w_dn, float *__restrict__ mo) {
    __shared__ float dist[64];
    const long bi = blockIdx.x;
    const int i = (int)(bi % N);
    if (threadIdx.x < I) {
        const float *ri = r + bi * 3;
        const int J = threadIdx.x;
        const float dx = ri[0] - R[J * 3], dy = ri[1] - R[J * 3 + 1], dz = ri[2] - R[J * 3 + 2];
        dist[J] = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    __syncthreads();
    const float *spa = i < U ? spa_up : spa_dn;
    const float *wt = i < U ? w_up : w_dn;
    float *p = mo + bi * (long)cols;
    for (int col = threadIdx.x; col < cols; col += blockDim.x) {
        float env = 0.f;
        for (int J = 0; J < I; ++J) env = __fadd_rn(env, __fmul_rn(wt[J * cols + col], expf(-spa[J * cols + col] * dist[J])));
        p[col] *= env;
    }
}

int launch_envelope(dpe_model *m, const float *r, int Bc, int C, float *mo, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    int cols = d.n_dets * d.n_el;
    if (C == 1 && d.n_ion <= 64) {
        k_envelope_fwd<<<Bc * d.n_el, 128, 0, s>>>(r, m->R_dev, d.n_el, d.n_up, d.n_ion, cols, m->sp_alpha[0], m->sp_alpha[1], m->env_w[0], m->env_w[1], mo);
        DPE_LAUNCH_CHECK(m);
        return DPE_OK;
    }
    long total = (long)Bc * d.n_el * cols;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    k_envelope<<<(int)blocks, 256, 0, s>>>(r, m->R_dev, Bc, d.n_el, d.n_up, d.n_ion, C, cols, m->sp_alpha[0], m->sp_alpha[1],
                                           m->env_w[0], m->env_w[1], mo);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// transferable atomic orbitals (orbitals/transferable_atomic_orbitals.py:287-349, cache branch; defaults
// use_el_ion_embedding = False, use_separate_ion_sum_for_envelopes = False, use_exponentials = True, full_det):
//   mo[i, d, k] = sum_J (h_i . b[J, k, d]) exp(-x[J, k, s(i, k), d] |r_i - R_J|)        k over [up orbitals | dn orbitals],
//   s = 0 if electron i and orbital k have the same spin, else 1 (:316-345)
// k_tao_pack lays the cache out for the kernels: tao_w[a][J*cols + d*N + k] = backflows_spin(k)[J, k_s, 0, d, a] (slice 0 for both
// spin types, :255-260), tao_ex[s][J*cols + d*N + k] = exponents_spin(k)[J, k_s, s, d].
// ------------------------------------------------------------------------------------------------
__global__ void k_tao_pack(const float *__restrict__ bf_up, const float *__restrict__ bf_dn, const float *__restrict__ ex_up,
                           const float *__restrict__ ex_dn, int I, int N, int U, int nd, int dl, float *__restrict__ w,
                           float *__restrict__ ex_same, float *__restrict__ ex_diff) {
    const int cols = nd * N;
    const long gc = (long)I * cols, total = (long)(dl + 1) * gc;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long a = idx / gc;                 // embedding feature, or dl for the exponent rows
        const long c = idx - a * gc;
        const int J = (int)(c / cols), dk = (int)(c - (long)J * cols), d = dk / N, k = dk - d * N;
        const bool up = k < U;
        const int ks = up ? k : k - U, n_orb = up ? U : N - U;
        const long base = (((long)J * n_orb + ks) * 2) * nd;          // [J][ks][spin type 0][d = 0]
        if (a < dl) {
            w[a * gc + c] = (up ? bf_up : bf_dn)[(base + d) * dl + a];
        } else {
            const float *ex = up ? ex_up : ex_dn;
            ex_same[c] = ex[base + d];
            ex_diff[c] = ex[base + nd + d];
        }
    }
}

int launch_tao_pack(dpe_model *m, const float *bf_up, const float *bf_dn, const float *ex_up, const float *ex_dn, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const int dl = d.n_hidden_one_el[d.n_iterations - 1];
    long total = (long)(dl + 1) * d.n_ion * d.n_dets * d.n_el;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    k_tao_pack<<<(int)blocks, 256, 0, s>>>(bf_up, bf_dn, ex_up, ex_dn, d.n_ion, d.n_el, d.n_up, d.n_dets, dl, m->tao_w, m->tao_ex[0], m->tao_ex[1]);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// One thread per (walker, electron, det*N + orbital); ions outer (the envelope of an ion is computed once), channels inner;
// the sum over ions accumulates in `mo` (first ion writes).  Product rule with e = exp(-x d): grad_i e = -x e (r_i - R_J)/d,
// lap e = e (x^2 - 2 x / d); only the three own-electron tangent channels and the Laplacian channel get extra terms.
__global__ void __launch_bounds__(256) k_tao_orbitals(const float *__restrict__ r, const float *__restrict__ R, int Bc, int N, int U,
                                                       int I, int C, int cols, const float *__restrict__ ex_same,
                                                       const float *__restrict__ ex_diff, const float *__restrict__ g,
                                                       float *__restrict__ mo) {
    const long total = (long)Bc * N * cols;
    const long ldg = (long)I * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int col = idx % cols;
        const long bi = idx / cols;
        const int i = bi % N, k = col % N;
        const float *ex = ((i < U) == (k < U)) ? ex_same : ex_diff;
        const float *ri = r + bi * 3;
        float *p = mo + bi * (long)C * cols + col;
        const int ci = 1 + 3 * i;
        for (int J = 0; J < I; ++J) {
            const float dx = ri[0] - R[J * 3], dy = ri[1] - R[J * 3 + 1], dz = ri[2] - R[J * 3 + 2];
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            const float x = ex[(long)J * cols + col];
            const float e = expf(-x * d);
            const float *gp = g + bi * (long)C * ldg + (long)J * cols + col;
            if (C == 1) {
                const float o = __fmul_rn(gp[0], e);              // no FMA contraction: same bits as the value channel below
                p[0] = J ? __fadd_rn(p[0], o) : o;
                continue;
            }
            const float inv = 1.f / d, ge = -x * e * inv;
            const float e1x = ge * dx, e1y = ge * dy, e1z = ge * dz;
            const float el = e * (x * x - 2.f * x * inv);
            const float bf0 = gp[0];
            const float t0 = gp[(long)ci * ldg], t1 = gp[(long)(ci + 1) * ldg], t2 = gp[(long)(ci + 2) * ldg];
            const float lap_extra = el * bf0 + 2.f * (e1x * t0 + e1y * t1 + e1z * t2);
            constexpr int UB = 8;
            for (int c0 = 0; c0 < C; c0 += UB) {
                float v[UB], acc[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (c0 + u < C) { v[u] = gp[(long)(c0 + u) * ldg]; acc[u] = J ? p[(long)(c0 + u) * cols] : 0.f; }
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int c = c0 + u;
                    if (c < C) {
                        float o = __fmul_rn(v[u], e);
                        if (c >= ci && c < ci + 3) o += (c == ci ? e1x : (c == ci + 1 ? e1y : e1z)) * bf0;
                        if (c == C - 1) o += lap_extra;
                        p[(long)c * cols] = __fadd_rn(acc[u], o);
                    }
                }
            }
        }
    }
}

int launch_tao_orbitals(dpe_model *m, const float *r, int Bc, int C, const float *g, float *mo, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    int cols = d.n_dets * d.n_el;
    long total = (long)Bc * d.n_el * cols;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    k_tao_orbitals<<<(int)blocks, 256, 0, s>>>(r, m->R_dev, Bc, d.n_el, d.n_up, d.n_ion, C, cols, m->tao_ex[0], m->tao_ex[1], g, mo);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// determinants: one group of T threads per (walker, determinant)
//   forward:    sign, log|det|                                 (jnp.linalg.slogdet, wavefunction.py:69)
//   Laplacian:  A^-1, g_k = tr(A^-1 dA_k), lap = tr(A^-1 lapA) - sum_k tr((A^-1 dA_k)^2)
// det record: [logdet, sign, lap, g_0 .. g_{K-1}]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v);

template <int T>
__device__ __forceinline__ double group_sum_d(double v, double *red, int tid) {
    v = warp_sum_d(v);
    if (T == 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < T / 32; ++w) s += red[w];
    return s;
}

__device__ __forceinline__ float det_rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Pivot search of a 16-lane half warp in ONE redux instruction: key = [28 high bits of |a| as an ordered integer | 15 - row], so the maximum key is
// the largest magnitude (ties and near-ties within 2^-20: the lower row, any of them is an equally good pivot).  Lanes outside [p, N) pass key 0
// ... except that row p itself must win when the whole column is zero: its key is at least 15 - p + 1 > 0 only if p < 15; a singular matrix gives
// log 0 = -inf either way.
__device__ __forceinline__ int half_warp_pivot(float mag, int q, bool candidate, int half) {
    const unsigned key = candidate ? ((__float_as_uint(mag) & 0xfffffff0u) | (unsigned)(15 - q)) : 0u;
    const unsigned best = __reduce_max_sync(half ? 0xffff0000u : 0x0000ffffu, key);
    return 15 - (int)(best & 15u);
}

template <int T>
__device__ __forceinline__ void group_sync() {
    if (T == 32) __syncwarp(); else __syncthreads();
}

template <int T>
__device__ __forceinline__ float group_sum(float v, float *red, int tid) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (T == 32) return v;
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < T / 32; ++w) s += red[w];
    return s;
}

template <int T, bool LAP, bool FACTOR = false>      // FACTOR: factor-only instantiation (tangent stage compiled out), see k_det_warp
__global__ void __launch_bounds__(T, FACTOR ? 12 : (T == 64 ? 6 : 3)) k_det(int N, int C, int n_det, const float *__restrict__ mo, float *__restrict__ det,
                                                                  float *__restrict__ ainv_hi, float *__restrict__ ainv_lo, int NP) {
    // The factorisation runs in FP64 (O(N^3) against the O(3N * N^3) FP32 tangent stage).
    extern __shared__ double smd[];
    // Gauss-Jordan IN PLACE: the identity half of [A | I] is never stored (half the FP64 work and shared memory of the augmented form);
    // row interchanges are recorded and undone as column interchanges of the inverse, in reverse order, through the index map cidx.
    const int W = N;
    const int S = W + 1;                 // padded row stride
    const int NQ = (N + 7) & ~7;         // rows padded to 8 floats (float4 loads), zero filled
    double *aug = smd;                   // [N][S]; after the sweep the same memory holds dA[2][N][NQ] and the tile exchange buffer
    size_t aug_b = (size_t)N * S * sizeof(double);                // must match launch_det
    {
        const size_t nbt = NQ >> 3, alias_b = (2 * (size_t)N * NQ + nbt * nbt * 64) * sizeof(float);
        if (LAP && !FACTOR && alias_b > aug_b) aug_b = alias_b;  // (the factor-only instantiation has no tangent stage: no alias, no AinvT)
        aug_b = (aug_b + 15) & ~(size_t)15;
    }
    float *AinvT = reinterpret_cast<float *>(reinterpret_cast<char *>(smd) + aug_b);   // [N][NQ]  AinvT[i][pos(o)] = Ainv[o][i]
    float *red = AinvT + (LAP && !FACTOR ? N * NQ : 0);           // 8 floats
    double *redd = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(red + 8) + 7) & ~uintptr_t(7));   // 8 doubles
    double *colp = redd + 8;                                      // [N] pivot column of the current step
    int *perm = reinterpret_cast<int *>(colp + N);                // [N] row exchanged with row p at step p; afterwards cidx
    __shared__ int piv_row;
    const int tid = threadIdx.x;
    const long bd = blockIdx.x;
    const long b = bd / n_det;
    const int dt = (int)(bd - b * n_det);
    const int cols = n_det * N;
    const float *mob = mo + b * (long)N * C * cols + (long)dt * N;     // element (i, c, o): mob[(i*C + c)*cols + o]
    const int K = C - 2;

    for (int e = tid; e < N * W; e += T) {
        int i = e / W, o = e - i * W;
        aug[i * S + o] = (double)mob[((long)i * C) * cols + o];
    }
    __syncthreads();
    LogDetAcc logdet;
    float sign = 1.f;
    for (int p = 0; p < N; ++p) {
        // pivot search (first maximum, as LAPACK's idamax)
        if (tid < 32) {
            double best = -1.0; int bi = p;
            for (int i = p + tid; i < N; i += 32) {
                double v = fabs(aug[i * S + p]);
                if (v > best) { best = v; bi = i; }
            }
            for (int o = 16; o; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) piv_row = bi;
        }
        __syncthreads();
        const int pr = piv_row;
        if (pr != p) {
            for (int o = tid; o < W; o += T) {
                double t = aug[p * S + o]; aug[p * S + o] = aug[pr * S + o]; aug[pr * S + o] = t;
            }
            sign = -sign;
            __syncthreads();
        }
        const double piv = aug[p * S + p];
        logdet.mul(piv);
        if (piv < 0.0) sign = -sign;
        const double inv = 1.0 / piv;
        __syncthreads();
        if (LAP) {
            // in-place Gauss-Jordan step: stash column p, scale the pivot row (its own entry becomes 1 / pivot), eliminate the column from
            // every other row (thread = column, no divisions); column p itself receives -a_ip / pivot
            if (tid == 0) perm[p] = pr;
            for (int i = tid; i < N; i += T) colp[i] = aug[i * S + p];
            __syncthreads();
            for (int o = tid; o < N; o += T) aug[p * S + o] = o == p ? inv : aug[p * S + o] * inv;
            __syncthreads();
            for (int o = tid; o < N; o += T) {
                if (o == p) {
                    for (int i = 0; i < N; ++i)
                        if (i != p) aug[i * S + p] = -colp[i] * inv;
                } else {
                    const double rp = aug[p * S + o];
                    for (int i = 0; i < N; ++i)
                        if (i != p) aug[i * S + o] = fma(-colp[i], rp, aug[i * S + o]);
                }
            }
            __syncthreads();
        } else {
            for (int o = p + 1 + tid; o < N; o += T) {
                const double rp = aug[p * S + o] * inv;
                for (int i = p + 1; i < N; ++i) aug[i * S + o] = fma(-aug[i * S + p], rp, aug[i * S + o]);
            }
            __syncthreads();
        }
    }
    float *out = det + bd * (long)(LAP ? K + 3 : 2);
    if (tid == 0) { out[0] = (float)logdet.value(); out[1] = sign; }
    if (!LAP) return;
    // Ainv = M P_{N-1} ... P_0 (M: the in-place result for the row-interchanged matrix): column i of Ainv is column cidx[i] of M
    if (tid == 0) {
        int *cidx = perm;                          // built in place from a local copy of the interchange list
        int pv[64];
        for (int k = 0; k < N; ++k) pv[k] = perm[k];
        for (int k = 0; k < N; ++k) cidx[k] = k;
        for (int pp = N - 1; pp >= 0; --pp) { const int t = cidx[pp]; cidx[pp] = cidx[pv[pp]]; cidx[pv[pp]] = t; }
    }
    __syncthreads();
    const int *cidx = perm;

    if (FACTOR || ainv_hi) {     // factor-only mode: the tensor-core trace kernel (det_tc.cu) consumes AinvT[i][sh + q] = Ainv[q][i], tf32-split, zero padded
        float *oh = ainv_hi + bd * (long)NP * NP, *ol = ainv_lo + bd * (long)NP * NP;
        const int sh = (dt * N) & 3;     // the TMA box starts at the 16-byte aligned column below det * N
        for (int e = tid; e < NP * NP; e += T) {
            const int i = e / NP, q = e - i * NP - sh;
            const float v = (i < N && q >= 0 && q < N) ? (float)aug[q * S + cidx[i]] : 0.f;
            const float hi = det_rna_tf32(v);
            oh[e] = hi;
            ol[e] = det_rna_tf32(v - hi);
        }
        return;
    }
    if constexpr (!FACTOR) {
    // ---- tangent stage: P_k = Ainv dA_k in 8 x 8 register tiles (one tile per thread, nb x nb <= T tiles) -------------
    // Rows of AinvT / dA are stored permuted, [first halves of the 8-column chunks | second halves], so that the float4
    // loads of a quarter warp hit distinct banks; padded rows / columns are zero, which zeroes the padding of P.
    const int nb = (N + 7) >> 3, NPAD = nb * 8;
    auto pos = [nb](int o) { return ((o & 7) >> 2) * (4 * nb) + (o >> 3) * 4 + (o & 3); };
    for (int e = tid; e < N * NPAD; e += T) AinvT[e] = 0.f;
    __syncthreads();
    for (int e = tid; e < N * N; e += T) {
        int i = e / N, o = e - i * N;
        AinvT[i * NPAD + pos(o)] = (float)aug[o * S + cidx[i]];
    }
    __syncthreads();                       // aug is dead from here on: its memory becomes dA[2] (double buffer) and Tx
    float *dAb = reinterpret_cast<float *>(aug);           // [2][N][NPAD]
    float *Tx = dAb + 2 * N * NPAD;                        // [nb*nb][64] transposed tiles for tr(P^2)
    for (int e = tid; e < 2 * N * NPAD; e += T) dAb[e] = 0.f;
    // record = lap' = tr(Ainv lapA) - sum_k tr(P_k^2) + sum_k tr(P_k)^2 (FP64 sums of FP32 P entries, see k_det_warp)
    double part = 0.0;
    for (int e = tid; e < N * N; e += T) {
        int i = e / N, o = e - i * N;
        part = fma((double)AinvT[i * NPAD + pos(o)], (double)mob[((long)i * C + C - 1) * cols + o], part);
    }
    const double lap = group_sum_d<T>(part, redd, tid);
    // dA_k staging with cp.async: thread -> column o (coalesced rows of mo), all rows i
    const int o_ld = tid;                                  // T >= N is guaranteed by the launcher
    const int pos_ld = o_ld < N ? pos(o_ld) : 0;
    auto stage = [&](int k, int buf) {
        if (o_ld < N) {
            const float *src = mob + (long)(1 + k) * cols + o_ld;
            float *dst = dAb + buf * N * NPAD + pos_ld;
            for (int i = 0; i < N; ++i) {
                const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst + i * NPAD);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d32), "l"(src + (long)i * C * cols) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int n_items = nb * nb;
    const bool has_item = tid < n_items;
    const int rb = has_item ? tid / nb : 0, cq = has_item ? tid - rb * nb : 0;
    const int partner = cq * nb + rb;
    double tr2 = 0.0, sum_g2 = 0.0;
    __syncthreads();
    stage(0, 0);
    for (int k = 0; k < K; ++k) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();                                   // dA_k visible; Tx of the previous k consumed
        if (k + 1 < K) stage(k + 1, (k + 1) & 1);
        const float *dA = dAb + (k & 1) * N * NPAD;
        double gkd = 0.0;
        float acc[8][8];
        if (has_item) {
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c2 = 0; c2 < 8; ++c2) acc[a][c2] = 0.f;
            const float *ap = AinvT + 4 * rb, *xp = dA + 4 * cq;
            for (int i = 0; i < N; ++i) {
                const float4 al = *reinterpret_cast<const float4 *>(ap + i * NPAD), ah = *reinterpret_cast<const float4 *>(ap + i * NPAD + 4 * nb);
                const float4 xl = *reinterpret_cast<const float4 *>(xp + i * NPAD), xh = *reinterpret_cast<const float4 *>(xp + i * NPAD + 4 * nb);
                const float av[8] = {al.x, al.y, al.z, al.w, ah.x, ah.y, ah.z, ah.w};
                const float xv[8] = {xl.x, xl.y, xl.z, xl.w, xh.x, xh.y, xh.z, xh.w};
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) acc[a][c2] = fmaf(av[a], xv[c2], acc[a][c2]);
            }
            if (rb == cq) {
#pragma unroll
                for (int a = 0; a < 8; ++a) gkd += (double)acc[a][a];
            }
            float *tx = Tx + tid * 64;                      // Tx[item][b][a] = acc[a][b]
#pragma unroll
            for (int c2 = 0; c2 < 8; ++c2) {
                *reinterpret_cast<float4 *>(tx + c2 * 8) = make_float4(acc[0][c2], acc[1][c2], acc[2][c2], acc[3][c2]);
                *reinterpret_cast<float4 *>(tx + c2 * 8 + 4) = make_float4(acc[4][c2], acc[5][c2], acc[6][c2], acc[7][c2]);
            }
        }
        __syncthreads();
        if (has_item) {                                    // tr(P^2) = sum over tiles <T(rb,cq), T(cq,rb)^T>
            const float *px = Tx + partner * 64;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const float4 p0 = *reinterpret_cast<const float4 *>(px + a * 8), p1 = *reinterpret_cast<const float4 *>(px + a * 8 + 4);
                tr2 = fma((double)acc[a][0], (double)p0.x, tr2); tr2 = fma((double)acc[a][1], (double)p0.y, tr2);
                tr2 = fma((double)acc[a][2], (double)p0.z, tr2); tr2 = fma((double)acc[a][3], (double)p0.w, tr2);
                tr2 = fma((double)acc[a][4], (double)p1.x, tr2); tr2 = fma((double)acc[a][5], (double)p1.y, tr2);
                tr2 = fma((double)acc[a][6], (double)p1.z, tr2); tr2 = fma((double)acc[a][7], (double)p1.w, tr2);
            }
        }
        const float gk = (float)group_sum_d<T>(gkd, redd, tid);
        sum_g2 = fma((double)gk, (double)gk, sum_g2);
        if (tid == 0) out[3 + k] = gk;
    }
    tr2 = group_sum_d<T>(tr2, redd, tid);
    if (tid == 0) out[2] = (float)(lap + (sum_g2 - tr2));
    }
}

// ------------------------------------------------------------------------------------------------
// Small systems (N <= 16): one warp per (walker, determinant), 4 warps per block. Lane = column of the augmented
// matrix [A | I] during the FP64 Gauss-Jordan sweep (no index arithmetic, conflict-free rows of 32 doubles);
// tangent stage in 8-column strips with zero-padded 16-float rows (2 x LDS.128 per 8 FMAs).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// FACTOR = the factor-only instantiation used with the tensor-core trace kernel: the SIMT tangent stage is compiled out, which halves
// the registers; the Gauss-Jordan sweep is a latency-bound chain per warp, so the extra resident warps translate into throughput.
template <bool LAP, bool FACTOR = false>
__global__ void __launch_bounds__(128, (LAP && !FACTOR) ? 4 : 8) k_det_warp(int N, int C, int n_det, long n_mat, const float *__restrict__ mo,
                                                   float *__restrict__ det, float *__restrict__ ainv_hi,
                                                   float *__restrict__ ainv_lo, int NP) {
    extern __shared__ double smd[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long bd = blockIdx.x * 4L + wib;
    if (bd >= n_mat) return;
    const int per_warp_doubles = N * 32 + ((LAP && !ainv_hi) ? (N * 16 + 2 * (N * 16 + 16) + 2 * 272 + 8) / 2 + 2 : 0);
    double *aug = smd + (size_t)wib * per_warp_doubles;      // [N][32]
    float *Ainv = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(aug + N * 32) + 15) & ~uintptr_t(15));   // [N][16], float4 loads
    float *dA = Ainv + N * 16;                                // [2][N*16 + 16]
    float *P = dA + 2 * (N * 16 + 16);                        // [2][16][17]
    const long b = bd / n_det;
    const int dt = (int)(bd - b * n_det);
    const int cols = n_det * N, K = C - 2, W = LAP ? 2 * N : N;
    const float *mob = mo + b * (long)N * C * cols + (long)dt * N;   // element (i, c, o): mob[(i*C + c)*cols + o]

    for (int i = 0; i < N; ++i) {
        double v = 0.0;
        if (lane < N) v = (double)mob[((long)i * C) * cols + lane];
        else if (lane - N == i) v = 1.0;
        aug[i * 32 + lane] = v;
    }
    __syncwarp();
    LogDetAcc logdet;
    float sign = 1.f;
    for (int p = 0; p < N; ++p) {
        float best = (lane >= p && lane < N) ? fabsf((float)aug[lane * 32 + p]) : -1.f;
        int bi = lane;
        for (int o = 16; o; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (bi != p) {
            double t = aug[p * 32 + lane];
            aug[p * 32 + lane] = aug[bi * 32 + lane];
            aug[bi * 32 + lane] = t;
            sign = -sign;
            __syncwarp();
        }
        const double piv = aug[p * 32 + p];
        logdet.mul(piv);
        if (piv < 0.0) sign = -sign;
        __syncwarp();
        if (LAP) {
            const double rowp = aug[p * 32 + lane] / piv;
            if (lane != p) aug[p * 32 + lane] = rowp;          // column p is left as is: it is never read again
            for (int i = 0; i < N; ++i) {
                const double fct = aug[i * 32 + p];
                if (i != p && lane != p && lane < W) aug[i * 32 + lane] = fma(-fct, rowp, aug[i * 32 + lane]);
            }
        } else {
            const double rowp = aug[p * 32 + lane] / piv;
            for (int i = p + 1; i < N; ++i) {
                const double fct = aug[i * 32 + p];
                if (lane > p && lane < N) aug[i * 32 + lane] = fma(-fct, rowp, aug[i * 32 + lane]);
            }
        }
        __syncwarp();
    }
    float *out = det + bd * (long)(LAP ? K + 3 : 2);
    if (lane == 0) { out[0] = (float)logdet.value(); out[1] = sign; }
    if (!LAP) return;

    if (FACTOR || ainv_hi) {     // factor-only mode (see k_det)
        float *oh = ainv_hi + bd * (long)NP * NP, *ol = ainv_lo + bd * (long)NP * NP;
        const int sh = (dt * N) & 3;     // the TMA box starts at the 16-byte aligned column below det * N
        for (int e = lane; e < NP * NP; e += 32) {
            const int i = e / NP, q = e - i * NP - sh;
            const float v = (i < N && q >= 0 && q < N) ? (float)aug[q * 32 + N + i] : 0.f;
            const float hi = det_rna_tf32(v);
            oh[e] = hi;
            ol[e] = det_rna_tf32(v - hi);
        }
        return;
    }
    if constexpr (!FACTOR) {
    // ---- tangent stage: two tangent directions per warp (half-warps), lane = column q of P_k -------------------------
    // P_k[o][q] = sum_i Ainv[o][i] dA_k[i][q]: per i one conflict-free LDS of dA_k[i][q] and four broadcast LDS.128 of the
    // row Ainv[:, i] feed 16 FMAs (the earlier row-strip version was shared-memory-bandwidth bound, ncu 84 %).
    float *AinvT = Ainv;                       // [N][16]: AinvT[i][o] = Ainv[o][i], zero padded to 16 columns
    const int HS = N * 16 + 16;                // half-warp stride of the dA staging buffers (keeps the halves on distinct banks)
    float *dAh = dA;                           // [2][N][16] (+16)
    float *Ps = P;                             // [2][16][17]
    for (int e = lane; e < N * 16; e += 32) AinvT[e] = 0.f;
    for (int e = lane; e < 2 * HS; e += 32) dAh[e] = 0.f;
    __syncwarp();
    for (int i = 0; i < N; ++i)
        if (lane < N) AinvT[i * 16 + lane] = (float)aug[lane * 32 + N + i];
    __syncwarp();
    const int half = lane >> 4, q = lane & 15;
    // element slots of this half-warp's matrix: e = q + 16 s -> (i, o)
    constexpr int NS = 16;                     // 16 * 16 >= N * N for N <= 16
    int src_off[NS], dst_off[NS];
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        int e = q + 16 * sl;
        int i = e / N, o = e - i * N;
        src_off[sl] = e < N * N ? (i * C) * cols + o : -1;
        dst_off[sl] = half * HS + i * 16 + o;
    }
    // lap' = tr(Ainv lapA) - sum_k tr(P_k^2) + sum_k tr(P_k)^2: for an ill-conditioned matrix P_k is nearly rank one,
    // tr(P_k^2) ~ tr(P_k)^2, and the two sums cancel to many digits -- they are accumulated in FP64 from the same FP32 P
    // entries so that the cancellation is exact with respect to those entries.
    double part = 0.0;
    for (int e = lane; e < N * N; e += 32) {
        int i = e / N, o = e - i * N;
        part = fma((double)AinvT[i * 16 + o], (double)mob[((long)i * C + C - 1) * cols + o], part);
    }
    const double lap = warp_sum_d(part);
    double t2 = 0.0, sum_g2 = 0.0;
    const int n_it = (K + 1) >> 1;
    float ld[NS];                              // software pipeline: the next pair of tangent matrices travels during the FMAs
    {
        const int k = half;
        const float *mk = mob + (long)(1 + k) * cols;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) ld[sl] = (src_off[sl] >= 0 && k < K) ? mk[src_off[sl]] : 0.f;
    }
    for (int it = 0; it < n_it; ++it) {
        const int k = 2 * it + half;
        const bool kv = k < K;
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < NS; ++sl)
            if (src_off[sl] >= 0) dAh[dst_off[sl]] = ld[sl];
        if (it + 1 < n_it) {
            const int kn = k + 2;
            const float *mk = mob + (long)(1 + kn) * cols;
#pragma unroll
            for (int sl = 0; sl < NS; ++sl) ld[sl] = (src_off[sl] >= 0 && kn < K) ? mk[src_off[sl]] : 0.f;
        }
        __syncwarp();
        float acc[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) acc[o] = 0.f;
        const float *dcol = dAh + half * HS + q;
        for (int i = 0; i < N; ++i) {
            const float x = dcol[i * 16];
            const float4 a0 = *reinterpret_cast<const float4 *>(AinvT + i * 16), a1 = *reinterpret_cast<const float4 *>(AinvT + i * 16 + 4);
            const float4 a2 = *reinterpret_cast<const float4 *>(AinvT + i * 16 + 8), a3 = *reinterpret_cast<const float4 *>(AinvT + i * 16 + 12);
            acc[0] = fmaf(a0.x, x, acc[0]); acc[1] = fmaf(a0.y, x, acc[1]); acc[2] = fmaf(a0.z, x, acc[2]); acc[3] = fmaf(a0.w, x, acc[3]);
            acc[4] = fmaf(a1.x, x, acc[4]); acc[5] = fmaf(a1.y, x, acc[5]); acc[6] = fmaf(a1.z, x, acc[6]); acc[7] = fmaf(a1.w, x, acc[7]);
            acc[8] = fmaf(a2.x, x, acc[8]); acc[9] = fmaf(a2.y, x, acc[9]); acc[10] = fmaf(a2.z, x, acc[10]); acc[11] = fmaf(a2.w, x, acc[11]);
            acc[12] = fmaf(a3.x, x, acc[12]); acc[13] = fmaf(a3.y, x, acc[13]); acc[14] = fmaf(a3.z, x, acc[14]); acc[15] = fmaf(a3.w, x, acc[15]);
        }
        // column q of P_k sits in acc[0..N): publish it, then read row q for tr(P^2) = sum_{o,q} P[o][q] P[q][o]
        float diag = 0.f;
        float *pcol = Ps + half * 272 + q;
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            pcol[o * 17] = acc[o];
            if (o == q) diag = acc[o];
        }
        double gkd = (q < N && kv) ? (double)diag : 0.0;
        for (int sh = 8; sh; sh >>= 1) gkd += __shfl_xor_sync(0xffffffffu, gkd, sh);    // within the half-warp
        const float gk = (float)gkd;
        if (kv) sum_g2 = (q == 0) ? fma((double)gk, (double)gk, sum_g2) : sum_g2;
        if (q == 0 && kv) out[3 + k] = gk;
        __syncwarp();
        if (q < N && kv) {
            const float *prow = Ps + half * 272 + q * 17;
#pragma unroll
            for (int o = 0; o < 16; ++o) t2 = fma((double)acc[o], (double)prow[o], t2);
        }
    }
    t2 = warp_sum_d(t2);
    sum_g2 = warp_sum_d(sum_g2);
    if (lane == 0) out[2] = (float)(lap + (sum_g2 - t2));
    }
}

// Forward-only determinants of small systems (N <= 16; the Metropolis step): the LU sweep needs N <= 16 columns, so a warp carries TWO
// matrices, one per half-warp (columns 0-15 / 16-31 of the same [N][32] FP64 tile; all shuffles stay inside a half).  Same arithmetic
// as k_det_warp<false>, twice the matrices per latency-bound sweep.
__global__ void __launch_bounds__(128, 8) k_det_fwd_half(int N, int n_det, long n_mat, const float *__restrict__ mo, float *__restrict__ det) {
    extern __shared__ double smd[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, q = lane & 15;
    const long bd = (blockIdx.x * 4L + wib) * 2 + half;
    const bool valid = bd < n_mat;
    double *aug = smd + (size_t)wib * N * 32 + half * 16;          // this half's columns: aug[i * 32 + q]
    const long bb = valid ? bd : 0;
    const long b = bb / n_det;
    const int dt = (int)(bb - b * n_det);
    const int cols = n_det * N;
    const float *mob = mo + b * (long)N * cols + (long)dt * N;     // C = 1: element (i, o) = mob[i * cols + o]
    for (int i = 0; i < N; ++i)
        aug[i * 32 + q] = (q < N && valid) ? (double)mob[(long)i * cols + q] : (q == i ? 1.0 : 0.0);
    __syncwarp();
    LogDetAcc logdet;
    float sign = 1.f;
    for (int p = 0; p < N; ++p) {
        int bi = half_warp_pivot(fabsf((float)aug[q * 32 + p]), q, q >= p && q < N, half);
        if (bi < p || bi >= N) bi = p;                               // all-zero column: keep the row (the determinant is zero anyway)
        if (bi != p) {
            double t = aug[p * 32 + q];
            aug[p * 32 + q] = aug[bi * 32 + q];
            aug[bi * 32 + q] = t;
            sign = -sign;
        }
        __syncwarp();
        const double piv = aug[p * 32 + p];
        logdet.mul(piv);
        if (piv < 0.0) sign = -sign;
        __syncwarp();
        const double rowp = aug[p * 32 + q] * (1.0 / piv);           // one reciprocal, not a division per element
        for (int i = p + 1; i < N; ++i) {
            const double fct = aug[i * 32 + p];
            if (q > p && q < N) aug[i * 32 + q] = fma(-fct, rowp, aug[i * 32 + q]);
        }
        __syncwarp();
    }
    if (q == 0 && valid) { det[bd * 2] = (float)logdet.value(); det[bd * 2 + 1] = sign; }
}

// Factor-only stage of the Laplacian pass for N <= 16 (what k_det_warp<true, true> did with one matrix per warp and the augmented [A | I]):
// TWO matrices per warp (16 lanes each, lane = column) and Gauss-Jordan IN PLACE, so every lane works on a live column; the inverse leaves as
// the tf32-split, zero-padded, column-shifted AinvT tile the tensor-core trace kernel consumes (see k_det).
__global__ void __launch_bounds__(128, 8) k_det_factor_half(int N, int C, int n_det, long n_mat, const float *__restrict__ mo, float *__restrict__ det,
                                                             float *__restrict__ ainv_hi, float *__restrict__ ainv_lo, int NP) {
    extern __shared__ double smd[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, half = lane >> 4, q = lane & 15;
    const long bd = (blockIdx.x * 4L + wib) * 2 + half;
    const bool valid = bd < n_mat;
    double *warp_base = smd + (size_t)wib * (N * 32 + 32 + 16);
    double *aug = warp_base + half * 16;                           // this half's columns: aug[i * 32 + q]
    double *colp = warp_base + N * 32 + half * 16;                 // [16] pivot column of the current step
    int *cidx = reinterpret_cast<int *>(warp_base + N * 32 + 32) + half * 16;      // [16] interchange list, then the column index map
    const long bb = valid ? bd : 0;
    const long b = bb / n_det;
    const int dt = (int)(bb - b * n_det);
    const int cols = n_det * N, K = C - 2;
    const float *mob = mo + b * (long)N * C * cols + (long)dt * N;   // element (i, c, o): mob[(i*C + c)*cols + o]
    for (int i = 0; i < N; ++i)
        aug[i * 32 + q] = (q < N && valid) ? (double)mob[((long)i * C) * cols + q] : (q == i ? 1.0 : 0.0);
    __syncwarp();
    LogDetAcc logdet;
    float sign = 1.f;
    for (int p = 0; p < N; ++p) {
        int bi = half_warp_pivot(fabsf((float)aug[q * 32 + p]), q, q >= p && q < N, half);
        if (bi < p || bi >= N) bi = p;                               // all-zero column: keep the row (the determinant is zero anyway)
        if (q == 0) cidx[p] = bi;
        if (bi != p) {
            double t = aug[p * 32 + q];
            aug[p * 32 + q] = aug[bi * 32 + q];
            aug[bi * 32 + q] = t;
            sign = -sign;
        }
        __syncwarp();
        const double piv = aug[p * 32 + p];
        logdet.mul(piv);
        if (piv < 0.0) sign = -sign;
        const double inv = 1.0 / piv;
        colp[q] = q < N ? aug[q * 32 + p] : 0.0;                   // column p before it is overwritten
        __syncwarp();
        const double rp = q == p ? inv : aug[p * 32 + q] * inv;
        if (q < N) aug[p * 32 + q] = rp;
        for (int i = 0; i < N; ++i) {
            if (i == p || q >= N) continue;
            aug[i * 32 + q] = q == p ? -colp[i] * inv : fma(-colp[i], rp, aug[i * 32 + q]);
        }
        __syncwarp();
    }
    if (q == 0) {
        if (valid) { float *out = det + bd * (long)(K + 3); out[0] = (float)logdet.value(); out[1] = sign; }
        int pv[16];
        for (int k = 0; k < N; ++k) pv[k] = cidx[k];
        for (int k = 0; k < N; ++k) cidx[k] = k;
        for (int pp = N - 1; pp >= 0; --pp) { const int t = cidx[pp]; cidx[pp] = cidx[pv[pp]]; cidx[pv[pp]] = t; }
    }
    __syncwarp();
    if (!valid) return;
    float *oh = ainv_hi + bd * (long)NP * NP, *ol = ainv_lo + bd * (long)NP * NP;
    const int sh = (dt * N) & 3;             // the TMA box starts at the 16-byte aligned column below det * N
    for (int e = q; e < NP * NP; e += 16) {
        const int i = e / NP, qq = e - i * NP - sh;
        const float v = (i < N && qq >= 0 && qq < N) ? (float)aug[qq * 32 + cidx[i]] : 0.f;
        const float hi = det_rna_tf32(v);
        oh[e] = hi;
        ol[e] = det_rna_tf32(v - hi);
    }
}

int launch_det(dpe_model *m, int Bc, int C, const float *mo, float *det, float *ainv, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const int N = d.n_el;
    const bool lap = C > 1;
    const int NP = det_tc_pad(N, d.n_dets);
    const bool force_generic = m->det_flags & 1, force_simt = m->det_flags & 2;   // dpe_set_det_path
    // Laplacian mode on the tensor-core path: FP64 factorisation here, traces of (dA_k Ainv) and (dA_k Ainv)^2 in det_tc.cu
    const bool tc = lap && ainv && NP <= 64 && m->gemm_path == 1 && !force_simt;
    float *ah = tc ? ainv : nullptr, *al = tc ? ainv + (size_t)Bc * d.n_dets * NP * NP : nullptr;
    const size_t nq = (N + 7) & ~7, nb = nq / 8;
    size_t aug_bytes = (size_t)N * (N + 1) * sizeof(double);                      // in-place Gauss-Jordan: no identity half
    const size_t alias_bytes = (2 * (size_t)N * nq + nb * nb * 64) * sizeof(float);
    if (lap && aug_bytes < alias_bytes) aug_bytes = alias_bytes;
    aug_bytes = (aug_bytes + 15) & ~(size_t)15;
    size_t smem = aug_bytes + ((lap ? (size_t)N * nq : 0) + 16) * sizeof(float) + 10 * sizeof(double) + (size_t)N * (sizeof(double) + sizeof(int)) + 16;
    int blocks = Bc * d.n_dets;
    {
    StageTimer t_factor(m, ST_DET_FACTOR, s);
    if (N <= 16 && !force_generic) {
        const size_t per_warp = ((size_t)N * 32 + ((lap && !tc) ? ((size_t)N * 16 + 2 * ((size_t)N * 16 + 16) + 2 * 272 + 8) / 2 + 2 : 0)) * sizeof(double);
        const long n_mat = (long)blocks;
        static const bool factor_single = getenv("DPE_DET_FACTOR_SINGLE") != nullptr;       // debug: one matrix per warp, augmented form
        if (lap && tc && !factor_single)
            k_det_factor_half<<<(int)((n_mat + 7) / 8), 128, 4 * ((size_t)N * 32 + 32 + 16) * sizeof(double), s>>>(N, C, d.n_dets, n_mat, mo, det, ah, al, NP);
        else if (lap && tc) k_det_warp<true, true><<<(int)((n_mat + 3) / 4), 128, 4 * per_warp + 32, s>>>(N, C, d.n_dets, n_mat, mo, det, ah, al, NP);
        else if (lap) k_det_warp<true><<<(int)((n_mat + 3) / 4), 128, 4 * per_warp + 32, s>>>(N, C, d.n_dets, n_mat, mo, det, ah, al, NP);
        else {
            static const bool one_per_warp = getenv("DPE_DET_FWD_SINGLE") != nullptr;      // debug: one matrix per warp
            if (one_per_warp) k_det_warp<false><<<(int)((n_mat + 3) / 4), 128, 4 * per_warp + 32, s>>>(N, C, d.n_dets, n_mat, mo, det, nullptr, nullptr, NP);
            else k_det_fwd_half<<<(int)((n_mat + 7) / 8), 128, 4 * (size_t)N * 32 * sizeof(double), s>>>(N, d.n_dets, n_mat, mo, det);
        }
    } else {
        // 64 threads: one 8 x 8 tile of P per thread (N <= 64 -> at most 64 tiles) and one column of mo per thread
        if (smem > DPE_SMEM_OPTIN) return set_error(DPE_ERR_UNSUPPORTED, "det: %zu bytes of shared memory", smem);
        {
            int e;
            if ((e = opt_in_smem(m, KID_DET64, k_det<64, true>))) return e;
            if ((e = opt_in_smem(m, KID_DET64F, k_det<64, true, true>))) return e;
            if ((e = opt_in_smem(m, KID_DET128, k_det<128, false>))) return e;
        }
        // factor-only: just the N x (N + 1) FP64 matrix, the pivot column and the interchange list
        const size_t smem_f = (((size_t)N * (N + 1) * sizeof(double) + 15) & ~(size_t)15) + 16 * sizeof(float) + 10 * sizeof(double) +
                              (size_t)N * (sizeof(double) + sizeof(int)) + 16;
        if (lap && tc) k_det<64, true, true><<<blocks, 64, smem_f, s>>>(N, C, d.n_dets, mo, det, ah, al, NP);
        else if (lap) k_det<64, true><<<blocks, 64, smem, s>>>(N, C, d.n_dets, mo, det, ah, al, NP);
        else {
            // forward pass: at most N - 1 <= 63 columns are live per pivot -- two warps per matrix, and only the N x (N + 1) matrix in shared memory
            static const bool wide_fwd = getenv("DPE_DET_FWD_128") != nullptr;
            const size_t smem_fw = (((size_t)N * (N + 1) * sizeof(double) + 15) & ~(size_t)15) + 16 * sizeof(float) + 10 * sizeof(double) +
                                   (size_t)N * (sizeof(double) + sizeof(int)) + 16;
            if (wide_fwd) k_det<128, false><<<blocks, 128, smem, s>>>(N, C, d.n_dets, mo, det, nullptr, nullptr, NP);
            else k_det<64, false><<<blocks, 64, smem_fw, s>>>(N, C, d.n_dets, mo, det, nullptr, nullptr, NP);
        }
    }
    DPE_LAUNCH_CHECK(m);
    }
    if (tc) {
        StageTimer t_trace(m, ST_DET_TRACE, s);
        int e = launch_det_trace_tc(m, Bc, C, mo, ah, al, NP, det, s);
        if (e == DPE_ERR_UNSUPPORTED) {      // shape outside the tensor-core kernel: redo the whole stage on CUDA cores
            if (N <= 16 && !force_generic) {
                const size_t per_warp = ((size_t)N * 32 + ((size_t)N * 16 + 2 * ((size_t)N * 16 + 16) + 2 * 272 + 8) / 2 + 2) * sizeof(double);
                k_det_warp<true><<<(int)((blocks + 3) / 4), 128, 4 * per_warp + 32, s>>>(N, C, d.n_dets, (long)blocks, mo, det, nullptr, nullptr, NP);
            } else {
                k_det<64, true><<<blocks, 64, smem, s>>>(N, C, d.n_dets, mo, det, nullptr, nullptr, NP);
            }
            DPE_LAUNCH_CHECK(m);
        } else if (e) {
            return e;
        }
    }
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// signed log-sum-exp over determinants (wavefunction.py:77-83) with gradient / Laplacian, then
// E_kin = -1/2 (1/2 lap L + 1/4 |grad L|^2), L = log psi^2 (hamiltonian.py:216). One warp per walker.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_combine(int Bc, int n_det, int K, bool lap_mode, const float *__restrict__ det,
                                                  const float *__restrict__ epot, float *__restrict__ phase,
                                                  float *__restrict__ logpsi2, float *__restrict__ grad,
                                                  float *__restrict__ ekin, float *__restrict__ eloc,
                                                  float *__restrict__ epot_out) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= Bc) return;
    const int rec = lap_mode ? K + 3 : 2;
    const float *db = det + (long)b * n_det * rec;
    // shift = max_d logdet, first arg-max
    float best = -INFINITY; int bi = 0;
    for (int d = lane; d < n_det; d += 32) {
        float v = db[(long)d * rec];
        if (v > best) { best = v; bi = d; }
    }
    for (int o = 16; o; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const float shift = best; const int md = bi;
    float psi = 0.f;
    for (int d = lane; d < n_det; d += 32) psi += db[(long)d * rec + 1] * expf(db[(long)d * rec] - shift);
    for (int o = 16; o; o >>= 1) psi += __shfl_xor_sync(0xffffffffu, psi, o);
    const float apsi = fabsf(psi);
    const float lp2 = 2.f * (logf(apsi + 1e-8f) + shift);
    if (lane == 0) {
        logpsi2[b] = lp2;
        if (phase) phase[b] = psi < 0.f ? 3.14159265358979323846f : 0.f;
    }
    if (!lap_mode) return;
    // The determinant record holds lap'_d = lap_d + sum_k g_dk^2; the sums over k and d below cancel to many digits when
    // one determinant dominates, so they run in FP64 on the FP32 records.
    const double inv_psi = 1.0 / (double)psi;
    const double rho = (double)apsi / ((double)apsi + 1e-8);
    double wl = 0.0;
    for (int d = lane; d < n_det; d += 32) {
        double w = (double)(db[(long)d * rec + 1] * expf(db[(long)d * rec] - shift)) * inv_psi;
        wl = fma(w, (double)db[(long)d * rec + 2], wl);
    }
    wl = warp_sum_d(wl);
    double sum_G2 = 0.0, sum_dev2 = 0.0, sum_f2 = 0.0, sum_gm2 = 0.0;
    for (int k = lane; k < K; k += 32) {
        double G = 0.0;
        for (int d = 0; d < n_det; ++d) {
            double w = (double)(db[(long)d * rec + 1] * expf(db[(long)d * rec] - shift)) * inv_psi;
            G = fma(w, (double)db[(long)d * rec + 3 + k], G);
        }
        const double gm = (double)db[(long)md * rec + 3 + k];
        const double Fk = rho * (G - gm) + gm;
        sum_G2 = fma(G, G, sum_G2);
        sum_gm2 = fma(gm, gm, sum_gm2);
        sum_dev2 = fma(G - gm, G - gm, sum_dev2);
        const float gr = (float)(2.0 * Fk);
        sum_f2 = fma((double)gr, (double)gr, sum_f2);
        if (grad) grad[(long)b * K + k] = gr;
    }
    sum_G2 = warp_sum_d(sum_G2); sum_gm2 = warp_sum_d(sum_gm2); sum_dev2 = warp_sum_d(sum_dev2); sum_f2 = warp_sum_d(sum_f2);
    if (lane == 0) {
        const double Gkk = wl - sum_G2;                                  // sum_k d_kk log|psi| without the epsilon
        const double l_m = (double)db[(long)md * rec + 2] - sum_gm2;      // Laplacian of the arg-max determinant alone
        const double F_lap = rho * (1.0 - rho) * sum_dev2 + rho * (Gkk - l_m) + l_m;
        const double lapL = 2.0 * F_lap;
        const float ek = (float)(-0.5 * (0.5 * lapL + 0.25 * sum_f2));
        if (ekin) ekin[b] = ek;
        if (epot_out) epot_out[b] = epot[b];
        if (eloc) eloc[b] = ek + epot[b];
    }
}

int launch_combine(dpe_model *m, int Bc, int C, const float *det, const float *epot, float *phase, float *logpsi2,
                   float *grad, float *ekin, float *eloc, float *epot_out, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    k_combine<<<(Bc * 32 + 127) / 128, 128, 0, s>>>(Bc, d.n_dets, 3 * d.n_el, C > 1, det, epot, phase, logpsi2, grad, ekin,
                                                    eloc, epot_out);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

}  // namespace dpe
