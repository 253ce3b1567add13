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

int launch_envelope(dpe_model *m, const float *r, int Bc, int C, float *mo, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    int cols = d.n_dets * d.n_el;
    long total = (long)Bc * d.n_el * cols;
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    k_envelope<<<(int)blocks, 256, 0, s>>>(r, m->R_dev, Bc, d.n_el, d.n_up, d.n_ion, C, cols, m->sp_alpha[0], m->sp_alpha[1],
                                           m->env_w[0], m->env_w[1], mo);
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

template <int T, bool LAP>
__global__ void __launch_bounds__(T, 3) k_det(int N, int C, int n_det, const float *__restrict__ mo, float *__restrict__ det) {
    // The factorisation runs in FP64 (O(N^3) against the O(3N * N^3) FP32 tangent stage).
    extern __shared__ double smd[];
    const int W = LAP ? 2 * N : N;       // augmented width
    const int S = W + 1;                 // padded row stride
    const int NP2 = (N + 1) & ~1;        // rows of Ainv^T padded to an even count (float2 loads)
    const int NQ = (N + 7) & ~7;         // tangent-matrix rows padded to 8 floats (float4 loads), zero filled
    double *aug = smd;                   // [N][S]; after the sweep the same memory holds dA [N][NQ] and P [N][NQ]
    float *AinvT = reinterpret_cast<float *>(aug + N * S);        // [N][NP2]  AinvT[i][o] = Ainv[o][i]
    float *red = AinvT + (LAP ? N * NP2 : 0);                     // 8 floats
    double *redd = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(red + 8) + 7) & ~uintptr_t(7));   // 8 doubles
    float *dA = reinterpret_cast<float *>(aug);
    float *P = dA + N * NQ;
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
        aug[i * S + o] = o < N ? (double)mob[((long)i * C) * cols + o] : ((o - N == i) ? 1.0 : 0.0);
    }
    __syncthreads();
    double logdet = 0.0;
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
        logdet += log(fabs(piv));
        if (piv < 0.0) sign = -sign;
        const double inv = 1.0 / piv;
        __syncthreads();
        if (LAP) {
            // Gauss-Jordan: scale the pivot row, eliminate the column from every other row (thread = column, no divisions)
            for (int o = tid; o < W; o += T) aug[p * S + o] *= inv;
            __syncthreads();
            for (int o = tid; o < W; o += T) {
                if (o == p) continue;
                const double rp = aug[p * S + o];
                for (int i = 0; i < N; ++i)
                    if (i != p) aug[i * S + o] = fma(-aug[i * S + p], rp, aug[i * S + o]);
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
    if (tid == 0) { out[0] = (float)logdet; out[1] = sign; }
    if (!LAP) return;

    for (int e = tid; e < N * NP2; e += T) {
        int i = e / NP2, o = e - i * NP2;
        AinvT[e] = o < N ? (float)aug[o * S + N + i] : 0.f;
    }
    __syncthreads();                       // aug is dead from here on: its memory becomes dA / P
    for (int e = tid; e < 2 * N * NQ; e += T) dA[e] = 0.f;
    // element -> (row, column) maps hoisted out of the k loop (runtime N: integer divisions are expensive)
    constexpr int ME = T == 32 ? 8 : 16;   // register-cached element slots per thread (covers N*N <= ME*T)
    int src_off[ME], dst_off[ME];
#pragma unroll
    for (int sl = 0; sl < ME; ++sl) {
        int e = tid + sl * T;
        int i = e / N, o = e - i * N;
        src_off[sl] = e < N * N ? (i * C) * cols + o : -1;
        dst_off[sl] = i * NQ + o;
    }
    const bool cached = N * N <= ME * T;
    // record = lap' = tr(Ainv lapA) - sum_k tr(P_k^2) + sum_k tr(P_k)^2 (FP64 sums of FP32 P entries, see k_det_warp)
    double part = 0.0;
    for (int e = tid; e < N * N; e += T) {
        int i = e / N, o = e - i * N;
        part = fma((double)AinvT[i * NP2 + o], (double)mob[((long)i * C + C - 1) * cols + o], part);
    }
    const double lap = group_sum_d<T>(part, redd, tid);
    // P = Ainv dA_k in 2 x 8 register tiles: per i one float2 (two rows of Ainv) and two float4 (eight columns of dA) for 16 FMAs
    const int nqc = NQ >> 3, n_items = (NP2 >> 1) * nqc;
    double tr2 = 0.0, sum_g2 = 0.0;
    float pre[ME];                                   // software pipeline: dA_{k+1} travels while P_k is computed
    if (cached) {
#pragma unroll
        for (int sl = 0; sl < ME; ++sl) pre[sl] = src_off[sl] >= 0 ? mob[(long)cols + src_off[sl]] : 0.f;
    }
    for (int k = 0; k < K; ++k) {
        const float *mk = mob + (long)(1 + k) * cols;
        __syncthreads();
        if (cached) {
#pragma unroll
            for (int sl = 0; sl < ME; ++sl)
                if (src_off[sl] >= 0) dA[dst_off[sl]] = pre[sl];
            if (k + 1 < K) {
#pragma unroll
                for (int sl = 0; sl < ME; ++sl)
                    if (src_off[sl] >= 0) pre[sl] = mk[(long)cols + src_off[sl]];
            }
        } else {
            for (int e = tid; e < N * N; e += T) {
                int i = e / N, o = e - i * N;
                dA[i * NQ + o] = mk[((long)i * C) * cols + o];
            }
        }
        __syncthreads();
        double gkd = 0.0;
        for (int item = tid; item < n_items; item += T) {
            const int o0 = (item / nqc) * 2, q0 = (item - (item / nqc) * nqc) * 8;
            float a0[8], a1[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) { a0[t] = 0.f; a1[t] = 0.f; }
            for (int i = 0; i < N; ++i) {
                const float2 av = *reinterpret_cast<const float2 *>(AinvT + i * NP2 + o0);
                const float4 x0 = *reinterpret_cast<const float4 *>(dA + i * NQ + q0);
                const float4 x1 = *reinterpret_cast<const float4 *>(dA + i * NQ + q0 + 4);
                a0[0] = fmaf(av.x, x0.x, a0[0]); a0[1] = fmaf(av.x, x0.y, a0[1]); a0[2] = fmaf(av.x, x0.z, a0[2]); a0[3] = fmaf(av.x, x0.w, a0[3]);
                a0[4] = fmaf(av.x, x1.x, a0[4]); a0[5] = fmaf(av.x, x1.y, a0[5]); a0[6] = fmaf(av.x, x1.z, a0[6]); a0[7] = fmaf(av.x, x1.w, a0[7]);
                a1[0] = fmaf(av.y, x0.x, a1[0]); a1[1] = fmaf(av.y, x0.y, a1[1]); a1[2] = fmaf(av.y, x0.z, a1[2]); a1[3] = fmaf(av.y, x0.w, a1[3]);
                a1[4] = fmaf(av.y, x1.x, a1[4]); a1[5] = fmaf(av.y, x1.y, a1[5]); a1[6] = fmaf(av.y, x1.z, a1[6]); a1[7] = fmaf(av.y, x1.w, a1[7]);
            }
            *reinterpret_cast<float4 *>(P + o0 * NQ + q0) = make_float4(a0[0], a0[1], a0[2], a0[3]);
            *reinterpret_cast<float4 *>(P + o0 * NQ + q0 + 4) = make_float4(a0[4], a0[5], a0[6], a0[7]);
            if (o0 + 1 < N) {
                *reinterpret_cast<float4 *>(P + (o0 + 1) * NQ + q0) = make_float4(a1[0], a1[1], a1[2], a1[3]);
                *reinterpret_cast<float4 *>(P + (o0 + 1) * NQ + q0 + 4) = make_float4(a1[4], a1[5], a1[6], a1[7]);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (q0 + t == o0) gkd += (double)a0[t];
                if (q0 + t == o0 + 1 && o0 + 1 < N) gkd += (double)a1[t];
            }
        }
        __syncthreads();
        for (int item = tid; item < n_items; item += T) {     // tr(P^2) = sum_{o,q} P[o][q] P[q][o] over this thread's tile
            const int o0 = (item / nqc) * 2, q0 = (item - (item / nqc) * nqc) * 8;
            const bool two = o0 + 1 < N;
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (q0 + t < N) {
                    tr2 = fma((double)P[o0 * NQ + q0 + t], (double)P[(q0 + t) * NQ + o0], tr2);
                    if (two) tr2 = fma((double)P[(o0 + 1) * NQ + q0 + t], (double)P[(q0 + t) * NQ + o0 + 1], tr2);
                }
        }
        const float gk = (float)group_sum_d<T>(gkd, redd, tid);
        sum_g2 = fma((double)gk, (double)gk, sum_g2);
        if (tid == 0) out[3 + k] = gk;
    }
    tr2 = group_sum_d<T>(tr2, redd, tid);
    if (tid == 0) out[2] = (float)(lap + (sum_g2 - tr2));
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

template <bool LAP>
__global__ void __launch_bounds__(128, 4) k_det_warp(int N, int C, int n_det, long n_mat, const float *__restrict__ mo,
                                                   float *__restrict__ det) {
    extern __shared__ double smd[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long bd = blockIdx.x * 4L + wib;
    if (bd >= n_mat) return;
    const int per_warp_doubles = N * 32 + (LAP ? (N * (N + 1) + N * 16 + N * 17 + 1) / 2 + 2 : 0);
    double *aug = smd + (size_t)wib * per_warp_doubles;      // [N][32]
    float *Ainv = reinterpret_cast<float *>(aug + N * 32);    // [N][N+1]
    float *dA = Ainv + N * (N + 1);                           // [N][16]
    float *P = dA + N * 16;                                   // [N][17]
    dA = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(dA) + 15) & ~uintptr_t(15));   // float4 loads
    if (LAP) P = dA + N * 16;
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
    double logdet = 0.0;
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
        logdet += log(fabs(piv));
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
    if (lane == 0) { out[0] = (float)logdet; out[1] = sign; }
    if (!LAP) return;

    for (int o = 0; o < N; ++o)
        if (lane < N) Ainv[o * (N + 1) + lane] = (float)aug[o * 32 + N + lane];
    for (int e = lane; e < N * 16; e += 32) dA[e] = 0.f;       // zero padding of the 16-float rows
    // element slots (hoisted index math): e = lane + 32 s -> (i, o)
    int src_off[8], d16[8], p17[8], t17[8];
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
        int e = lane + 32 * sl;
        int i = e / N, o = e - i * N;
        bool ok = e < N * N;
        src_off[sl] = ok ? (i * C) * cols + o : -1;
        d16[sl] = i * 16 + o;
        p17[sl] = i * 17 + o;
        t17[sl] = o * 17 + i;
    }
    __syncwarp();
    // The record holds lap' = tr(Ainv lapA) - sum_k tr(P_k^2) + sum_k tr(P_k)^2: for an ill-conditioned matrix P_k is
    // nearly rank one, tr(P_k^2) ~ tr(P_k)^2, and the two sums cancel to many digits -- they are accumulated in FP64
    // from the same FP32 P entries so that the cancellation is exact with respect to those entries.
    double part = 0.0;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl)
        if (src_off[sl] >= 0) {
            int e = lane + 32 * sl, i = e / N, o = e - i * N;
            part = fma((double)Ainv[o * (N + 1) + i], (double)mob[(long)(C - 1) * cols + src_off[sl]], part);
        }
    const double lap = warp_sum_d(part);
    const int orow = lane >> 1, q0 = (lane & 1) * 8;
    const bool strip = orow < N;
    double t2 = 0.0, sum_g2 = 0.0;
    float ld[8];                                     // software pipeline: dA_{k+1} travels while P_k is computed
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) ld[sl] = src_off[sl] >= 0 ? mob[(long)cols + src_off[sl]] : 0.f;
    for (int k = 0; k < K; ++k) {
        const float *mk = mob + (long)(1 + k) * cols;
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < 8; ++sl)
            if (src_off[sl] >= 0) dA[d16[sl]] = ld[sl];
        if (k + 1 < K) {
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
                if (src_off[sl] >= 0) ld[sl] = mk[(long)cols + src_off[sl]];
        }
        __syncwarp();
        double gkd = 0.0;
        if (strip) {
            float acc[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[t] = 0.f;
            const float *arow = Ainv + orow * (N + 1);
            for (int i = 0; i < N; ++i) {
                const float av = arow[i];
                const float4 x0 = *reinterpret_cast<const float4 *>(dA + i * 16 + q0);
                const float4 x1 = *reinterpret_cast<const float4 *>(dA + i * 16 + q0 + 4);
                acc[0] = fmaf(av, x0.x, acc[0]); acc[1] = fmaf(av, x0.y, acc[1]); acc[2] = fmaf(av, x0.z, acc[2]); acc[3] = fmaf(av, x0.w, acc[3]);
                acc[4] = fmaf(av, x1.x, acc[4]); acc[5] = fmaf(av, x1.y, acc[5]); acc[6] = fmaf(av, x1.z, acc[6]); acc[7] = fmaf(av, x1.w, acc[7]);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                P[orow * 17 + q0 + t] = acc[t];
                if (q0 + t == orow) gkd = (double)acc[t];
            }
        }
        const float gk = (float)warp_sum_d(gkd);
        sum_g2 = fma((double)gk, (double)gk, sum_g2);       // with the rounded value that the record stores
        if (lane == 0) out[3 + k] = gk;
        __syncwarp();
#pragma unroll
        for (int sl = 0; sl < 8; ++sl)
            if (src_off[sl] >= 0) t2 = fma((double)P[p17[sl]], (double)P[t17[sl]], t2);
    }
    t2 = warp_sum_d(t2);
    if (lane == 0) out[2] = (float)(lap + (sum_g2 - t2));
}

int launch_det(dpe_model *m, int Bc, int C, const float *mo, float *det, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const int N = d.n_el;
    const bool lap = C > 1;
    const size_t np2 = (N + 1) & ~1, nq = (N + 7) & ~7;
    size_t aug_bytes = (size_t)N * ((lap ? 2 * N : N) + 1) * sizeof(double);
    if (lap && aug_bytes < 2 * N * nq * sizeof(float)) aug_bytes = 2 * N * nq * sizeof(float);
    size_t smem = aug_bytes + ((lap ? (size_t)N * np2 : 0) + 16) * sizeof(float) + 10 * sizeof(double);
    int blocks = Bc * d.n_dets;
    static const bool force_generic = getenv("DPE_DET_GENERIC") != nullptr;   // debug knob
    if (N <= 16 && !force_generic) {
        const size_t per_warp = ((size_t)N * 32 + (lap ? ((size_t)N * (N + 1) + N * 16 + N * 17 + 1) / 2 + 2 : 0)) * sizeof(double);
        const long n_mat = (long)blocks;
        if (lap) k_det_warp<true><<<(int)((n_mat + 3) / 4), 128, 4 * per_warp + 32, s>>>(N, C, d.n_dets, n_mat, mo, det);
        else k_det_warp<false><<<(int)((n_mat + 3) / 4), 128, 4 * per_warp + 32, s>>>(N, C, d.n_dets, n_mat, mo, det);
    } else if (N <= 16) {
        if (lap) k_det<32, true><<<blocks, 32, smem, s>>>(N, C, d.n_dets, mo, det);
        else k_det<32, false><<<blocks, 32, smem, s>>>(N, C, d.n_dets, mo, det);
    } else {
        if (smem > 48 * 1024) {
            DPE_CUDA(cudaFuncSetAttribute(k_det<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            DPE_CUDA(cudaFuncSetAttribute(k_det<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        if (lap) k_det<128, true><<<blocks, 128, smem, s>>>(N, C, d.n_dets, mo, det);
        else k_det<128, false><<<blocks, 128, smem, s>>>(N, C, d.n_dets, mo, det);
    }
    DPE_LAUNCH_CHECK(m);
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
