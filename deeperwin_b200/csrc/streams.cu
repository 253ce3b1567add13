// Feature construction, el-ion stream, el-el pair stream, SchNet convolution, spin means and the
// tanh value/tangent/Laplacian rule of the dpe4 embedding (reference: model/input_features.py:112-303,
// model/embeddings/ferminet_embedding.py:45-267, model/mlp.py:45-69, utils/utils.py:262-305).
//
// Layout (walker-major, feature-minor so that warps read/write 128-byte rows):
//   one-electron stream  X[b][i][c][ld]   columns: h_one (d_in) | conv_ee (emb) | conv_eI (dE)
//   channels c: 0 = value, 1+3e+a = d/d r_{e,a}, 3N+1 = Laplacian   (C = 3N+2; C = 1 forward-only)
//   el-ion stream        5 channels (value, d/d r_i{x,y,z}, Laplacian): depends on r_i only
//   pair stream          3 channels (f, f', f'') as a function of the scalar distance d_ij
#include "dpe_internal.cuh"

namespace dpe {

__device__ __forceinline__ float tanh_f32(float x) { return tanhf(x); }

// ------------------------------------------------------------------------------------------------
// features: h_one^0 = reshape([dist_eI, diff_eI]) with tangents, plus E_pot
// ------------------------------------------------------------------------------------------------
// Forward pass (one channel): a handful of values per electron row, one thread per value.
__global__ void k_features_fwd(const float *__restrict__ r, const float *__restrict__ R, int n_rows, int I, float *__restrict__ x0, int ldx) {
    const int d0 = 4 * I;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * d0) return;
    const int row = idx / d0, col = idx - row * d0, J = col >> 2, q = col & 3;
    const float *ri = r + (long)row * 3;
    const float dx = ri[0] - R[J * 3 + 0], dy = ri[1] - R[J * 3 + 1], dz = ri[2] - R[J * 3 + 2];
    x0[(long)row * ldx + col] = q == 0 ? sqrtf(dx * dx + dy * dy + dz * dz) : (q == 1 ? dx : (q == 2 ? dy : dz));
}

// One block per (walker, electron): the el-ion distances and differences once per block, then the C x 4I entries of the electron's rows with
// 32-bit index arithmetic (the first version spent ~100 instructions of 64-bit divisions per 4-byte store).
__global__ void __launch_bounds__(128) k_features(const float *__restrict__ r, const float *__restrict__ R, int N, int I, int C,
                                                   float *__restrict__ x0, int ldx) {
    extern __shared__ float fs[];                  // [I][4]: d, dx, dy, dz
    const long bi = blockIdx.x;                    // b * N + i
    const int i = (int)(bi % N);
    for (int J = threadIdx.x; J < I; J += blockDim.x) {
        const float *ri = r + bi * 3;
        const float dx = ri[0] - R[J * 3 + 0], dy = ri[1] - R[J * 3 + 1], dz = ri[2] - R[J * 3 + 2];
        fs[J * 4] = sqrtf(dx * dx + dy * dy + dz * dz);
        fs[J * 4 + 1] = dx; fs[J * 4 + 2] = dy; fs[J * 4 + 3] = dz;
    }
    __syncthreads();
    const int d0 = 4 * I, own = 1 + 3 * i;
    float *xb = x0 + bi * (long)C * ldx;
    for (int t = threadIdx.x; t < C * d0; t += blockDim.x) {
        const int c = t / d0, col = t - c * d0, J = col >> 2, q = col & 3;
        const float d = fs[J * 4];
        float v = 0.f;
        if (c == 0) v = fs[J * 4 + q];                                  // (d, dx, dy, dz)
        else if (c == C - 1) v = q == 0 ? 2.0f / d : 0.f;               // Laplacian of |r - R|
        else if (c >= own && c < own + 3) {                              // tangent of this electron's own coordinate a
            const int a = c - own;
            v = q == 0 ? fs[J * 4 + 1 + a] / d : (q - 1 == a ? 1.f : 0.f);
        }
        xb[(long)c * ldx + col] = v;
    }
}

// E_pot (hamiltonian.py:17-39): one warp per walker.
__global__ void k_epot(const float *__restrict__ r, const float *__restrict__ R, const float *__restrict__ Zf, int Bc,
                       int N, int I, const float *__restrict__ e_ion_ion, float *__restrict__ epot) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= Bc) return;
    const float *rb = r + (long)warp * N * 3;
    // FP32 terms (as the reference computes them), summed in FP64: benzene has 1365 of them and an FP32 running sum would be the
    // largest error of the whole potential energy
    double e_ei = 0.0, e_ee = 0.0;
    for (int t = lane; t < N * I; t += 32) {
        int i = t / I, J = t - i * I;
        float dx = rb[i * 3] - R[J * 3], dy = rb[i * 3 + 1] - R[J * 3 + 1], dz = rb[i * 3 + 2] - R[J * 3 + 2];
        e_ei += (double)(Zf[J] / sqrtf(dx * dx + dy * dy + dz * dz));
    }
    for (int t = lane; t < N * N; t += 32) {
        int i = t / N, j = t - i * N;
        if (j > i) {
            float dx = rb[i * 3] - rb[j * 3], dy = rb[i * 3 + 1] - rb[j * 3 + 1], dz = rb[i * 3 + 2] - rb[j * 3 + 2];
            e_ee += (double)(1.0f / sqrtf(dx * dx + dy * dy + dz * dz));
        }
    }
    for (int o = 16; o; o >>= 1) {
        e_ei += __shfl_xor_sync(0xffffffffu, e_ei, o);
        e_ee += __shfl_xor_sync(0xffffffffu, e_ee, o);
    }
    if (lane == 0) epot[warp] = (float)(e_ee - e_ei + (double)e_ion_ion[0]);
}

int launch_features(dpe_model *m, const float *r, int Bc, int C, float *x0, int ldx, float *epot, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    if (C == 1 && (long)Bc * d.n_el * 4 * d.n_ion < (1L << 31)) {
        const int n = Bc * d.n_el * 4 * d.n_ion;
        k_features_fwd<<<(n + 255) / 256, 256, 0, s>>>(r, m->R_dev, Bc * d.n_el, d.n_ion, x0, ldx);
    } else {
        k_features<<<Bc * d.n_el, 128, (size_t)d.n_ion * 4 * sizeof(float), s>>>(r, m->R_dev, d.n_el, d.n_ion, C, x0, ldx);
    }
    DPE_LAUNCH_CHECK(m);
    if (epot) {
        k_epot<<<(Bc * 32 + 255) / 256, 256, 0, s>>>(r, m->R_dev, m->Z_dev, Bc, d.n_el, d.n_ion, m->eii_dev, epot);
        DPE_LAUNCH_CHECK(m);
    }
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// Tiny dense layer for one pair held by a warp: lane = output feature, the input vector x[c][0:din] sits in a
// per-warp shared buffer (float4 broadcast reads), weights are staged as [k/4][n][4] (one LDS.128 per 4 k).
template <int CH>
__device__ __forceinline__ void pair_dense_tanh(const float *__restrict__ W4, const float *__restrict__ bias, int din,
                                                int dout, const float *__restrict__ xs, float (&y)[CH], int lane) {
    float z[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) z[c] = 0.f;
    const bool act = lane < dout;
    if (din == 1) {
        const float w = act ? W4[lane * 4] : 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) z[c] = xs[c * 32] * w;
    } else {
        const int nk4 = din >> 2;
        for (int k4 = 0; k4 < nk4; ++k4) {
            const float4 w = act ? *reinterpret_cast<const float4 *>(W4 + (k4 * dout + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const float4 xv = *reinterpret_cast<const float4 *>(xs + c * 32 + k4 * 4);
                z[c] = fmaf(xv.x, w.x, z[c]); z[c] = fmaf(xv.y, w.y, z[c]); z[c] = fmaf(xv.z, w.z, z[c]); z[c] = fmaf(xv.w, w.w, z[c]);
            }
        }
    }
    if (act) z[0] += bias[lane];
    float t = tanh_f32(z[0]);
    float d1 = 1.f - t * t;
    y[0] = act ? t : 0.f;
    if (CH > 1) {
        float ssq = 0.f;
#pragma unroll
        for (int c = 1; c < CH - 1; ++c) {
            ssq = fmaf(z[c], z[c], ssq);
            y[c] = act ? d1 * z[c] : 0.f;
        }
        y[CH - 1] = act ? d1 * z[CH - 1] - 2.f * t * d1 * ssq : 0.f;
    }
}

struct SmallNet {         // shared-memory offsets (floats) of the per-iteration tiny layers
    int w_off[DPE_MAX_ITER][4], b_off[DPE_MAX_ITER][4];
    int din[DPE_MAX_ITER], dout_w[DPE_MAX_ITER], dout_h[DPE_MAX_ITER];
    int n_iter, total;
};

// ------------------------------------------------------------------------------------------------
// el-ion stream: h_eI^it for all iterations + conv_eI^it[b,i,:] = sum_J h_eI^it[b,i,J,:] * him^it[J,:]
// ------------------------------------------------------------------------------------------------
struct EionArgs {
    const float *r, *R;
    const float *w[DPE_MAX_ITER], *b[DPE_MAX_ITER], *him[DPE_MAX_ITER];
    float *out[DPE_MAX_ITER];
    int dE[DPE_MAX_ITER];
    int n_iter, N, I, n_rows;  // n_rows = Bc*N
};

template <int CH>
__global__ void __launch_bounds__(256) k_eion_stream(EionArgs a) {
    extern __shared__ __align__(16) float smem[];
    // stage the (n_iter-1) layer weights as [k/4][n][4] (+ bias): one LDS.128 feeds four k of a lane's output, and the layer input sits in a per-warp
    // shared buffer read with broadcast LDS.128 (pair_dense_tanh) -- the shuffle-per-k version was bound by the SM's one shuffle per clock
    int off = 0;
    int w_off[DPE_MAX_ITER], b_off[DPE_MAX_ITER];
    for (int it = 0; it + 1 < a.n_iter; ++it) {
        const int kp = (a.dE[it] + 3) & ~3;
        w_off[it] = off; off += kp * a.dE[it + 1];
        b_off[it] = off; off += (a.dE[it + 1] + 3) & ~3;
    }
    float *xs_all = smem + off;                  // [8 warps][CH][32]
    for (int it = 0; it + 1 < a.n_iter; ++it) {
        const int din = a.dE[it], dout = a.dE[it + 1], kp = (din + 3) & ~3;
        for (int t = threadIdx.x; t < kp * dout; t += blockDim.x) {
            const int k4 = t / (dout * 4), rem = t - k4 * dout * 4, n = rem >> 2, kk = rem & 3, k = k4 * 4 + kk;
            smem[w_off[it] + t] = k < din ? a.w[it][k * dout + n] : 0.f;
        }
        for (int t = threadIdx.x; t < dout; t += blockDim.x) smem[b_off[it] + t] = a.b[it][t];
    }
    for (int t = threadIdx.x; t < 8 * CH * 32; t += blockDim.x) xs_all[t] = 0.f;
    __syncthreads();
    float *xs = xs_all + (threadIdx.x >> 5) * CH * 32;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < a.n_rows; row += gridDim.x * warps_per_block) {
        const float *ri = a.r + (long)row * 3;
        float rx = ri[0], ry = ri[1], rz = ri[2];
        float acc[DPE_MAX_ITER][CH];
#pragma unroll
        for (int it = 0; it < DPE_MAX_ITER; ++it)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[it][c] = 0.f;
        for (int J = 0; J < a.I; ++J) {
            float dx = rx - a.R[J * 3], dy = ry - a.R[J * 3 + 1], dz = rz - a.R[J * 3 + 2];
            float d = sqrtf(dx * dx + dy * dy + dz * dz);
            float x[CH];
            {   // features [d, dx, dy, dz] on lanes 0..3
                float diff = lane == 1 ? dx : (lane == 2 ? dy : dz);
                x[0] = lane == 0 ? d : (lane < 4 ? diff : 0.f);
                if (CH > 1) {
                    float inv = 1.f / d;
                    float u[3] = {dx * inv, dy * inv, dz * inv};
#pragma unroll
                    for (int c = 1; c < CH - 1; ++c) x[c] = lane == 0 ? u[c - 1] : ((lane == c) ? 1.f : 0.f);
                    x[CH - 1] = lane == 0 ? 2.f * inv : 0.f;
                }
            }
#pragma unroll
            for (int it = 0; it < DPE_MAX_ITER; ++it) {
                if (it < a.n_iter) {
                    float hm = lane < a.dE[it] ? a.him[it][J * a.dE[it] + lane] : 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) acc[it][c] = fmaf(x[c], hm, acc[it][c]);
                    if (it + 1 < a.n_iter) {
                        float y[CH];
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < CH; ++c) xs[c * 32 + lane] = x[c];
                        __syncwarp();
                        pair_dense_tanh<CH>(smem + w_off[it], smem + b_off[it], (a.dE[it] + 3) & ~3, a.dE[it + 1], xs, y, lane);
                        const bool res = a.dE[it] == a.dE[it + 1];     // mlp.py:13-16
#pragma unroll
                        for (int c = 0; c < CH; ++c) x[c] = res ? (x[c] + y[c]) * 0.70710678118654752f : y[c];
                    }
                }
            }
        }
#pragma unroll
        for (int it = 0; it < DPE_MAX_ITER; ++it) {
            if (it < a.n_iter && lane < a.dE[it]) {
#pragma unroll
                for (int c = 0; c < CH; ++c) a.out[it][((long)row * CH + c) * a.dE[it] + lane] = acc[it][c];
            }
        }
    }
}

int launch_eion_stream(dpe_model *m, const float *r, int Bc, int CE, float *ei_base, const size_t *ei_off, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    EionArgs a;
    a.r = r; a.R = m->R_dev; a.n_iter = d.n_iterations; a.N = d.n_el; a.I = d.n_ion; a.n_rows = Bc * d.n_el;
    size_t smem = 0;
    for (int it = 0; it < d.n_iterations; ++it) {
        a.dE[it] = m->it[it].dE;
        a.him[it] = m->it[it].him;
        a.out[it] = ei_base + ei_off[it];
        a.w[it] = m->it[it].h_el_ion.w;
        a.b[it] = m->it[it].h_el_ion.b;
        if (it + 1 < d.n_iterations) smem += ((size_t)((m->it[it].dE + 3) & ~3) * m->it[it + 1].dE + ((m->it[it + 1].dE + 3) & ~3)) * sizeof(float);
    }
    smem += (size_t)8 * CE * 32 * sizeof(float);          // per-warp layer-input buffers
    int blocks = (a.n_rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (CE == 1) k_eion_stream<1><<<blocks, 256, smem, s>>>(a);
    else k_eion_stream<5><<<blocks, 256, smem, s>>>(a);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// pair stream: w^it[b,i,j,:] (value, d/dd, d2/dd2) for every iteration, one warp per pair
// ------------------------------------------------------------------------------------------------
struct PairArgs {
    const float *r;
    const float *ww[DPE_MAX_ITER][2], *wb[DPE_MAX_ITER][2];   // w_same / w_diff
    const float *hw[DPE_MAX_ITER][2], *hb[DPE_MAX_ITER][2];   // h_same / h_diff
    float *out[DPE_MAX_ITER];
    int dP[DPE_MAX_ITER];
    int n_iter, N, U, emb;
    long n_pairs;  // Bc*N*N
};

// One warp per UNORDERED electron pair (i <= j): w(i,j) = w(j,i) because both orderings see the same distance and the
// same (same-spin / different-spin) weights (ferminet_embedding.py:197-205: diff = concat(ud, du) through one MLP).
template <int CH>
__global__ void __launch_bounds__(256) k_pair_stream(PairArgs a) {
    extern __shared__ __align__(16) float smem[];
    // per iteration: [same block][diff block], each = w_w | w_b | (h_w | h_b); weights re-laid out as [k/4][n][4]
    int base[DPE_MAX_ITER], blk[DPE_MAX_ITER];
    int off = 0;
    for (int it = 0; it < a.n_iter; ++it) {
        base[it] = off;
        const int kp = a.dP[it] == 1 ? 4 : a.dP[it];        // din = 1 is padded to one k4 group
        blk[it] = kp * a.emb + a.emb + ((it + 1 < a.n_iter) ? kp * a.dP[it + 1] + a.dP[it + 1] : 0);
        off += 2 * blk[it];
    }
    float *xs_all = smem + off;                               // [8 warps][CH][32]
    int *pair_tab = reinterpret_cast<int *>(xs_all + 8 * CH * 32);   // [n_up_pairs] packed (i << 8) | j
    auto stage = [&](float *dst, const float *src, int din, int dout) {
        const int kp = din == 1 ? 4 : din;
        for (int t = threadIdx.x; t < kp * dout; t += blockDim.x) {
            int k4 = t / (dout * 4), rem = t - k4 * dout * 4, n = rem >> 2, kk = rem & 3, k = k4 * 4 + kk;
            dst[t] = k < din ? src[k * dout + n] : 0.f;
        }
    };
    for (int it = 0; it < a.n_iter; ++it)
        for (int sd = 0; sd < 2; ++sd) {
            float *dst = smem + base[it] + sd * blk[it];
            const int kp = a.dP[it] == 1 ? 4 : a.dP[it];
            stage(dst, a.ww[it][sd], a.dP[it], a.emb);
            for (int t = threadIdx.x; t < a.emb; t += blockDim.x) dst[kp * a.emb + t] = a.wb[it][sd][t];
            if (it + 1 < a.n_iter) {
                stage(dst + kp * a.emb + a.emb, a.hw[it][sd], a.dP[it], a.dP[it + 1]);
                for (int t = threadIdx.x; t < a.dP[it + 1]; t += blockDim.x) dst[kp * a.emb + a.emb + kp * a.dP[it + 1] + t] = a.hb[it][sd][t];
            }
        }
    const int n_up = a.N * (a.N + 1) / 2;
    for (int t = threadIdx.x; t < n_up; t += blockDim.x) {     // decode t -> (i <= j)
        int i = 0, rem = t;
        while (rem >= a.N - i) { rem -= a.N - i; ++i; }
        pair_tab[t] = (i << 8) | (i + rem);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *xs = xs_all + wib * CH * 32;
    const long n_units = a.n_pairs / (a.N * a.N) * n_up;       // walkers x unordered pairs
    for (long p = blockIdx.x * 8L + wib; p < n_units; p += gridDim.x * 8L) {
        const long b = p / n_up;
        const int pk = pair_tab[(int)(p - b * n_up)];
        const int i = pk >> 8, j = pk & 255;
        const float *ri = a.r + (b * a.N + i) * 3, *rj = a.r + (b * a.N + j) * 3;
        const float dx = rj[0] - ri[0], dy = rj[1] - ri[1], dz = rj[2] - ri[2];
        const float d = (i == j) ? 0.f : sqrtf(dx * dx + dy * dy + dz * dz);   // utils.py:299-300: diagonal exactly 0, zero grads
        const int sd = ((i < a.U) == (j < a.U)) ? 0 : 1;
        float x[CH];
        x[0] = lane == 0 ? d : 0.f;
        if (CH > 1) {
            x[1] = (lane == 0 && i != j) ? 1.f : 0.f;
            x[2] = 0.f;
        }
        const long o_ij = ((b * a.N + i) * a.N + j) * (long)CH * a.emb, o_ji = ((b * a.N + j) * a.N + i) * (long)CH * a.emb;
#pragma unroll
        for (int it = 0; it < DPE_MAX_ITER; ++it) {
            if (it < a.n_iter) {
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CH; ++c) xs[c * 32 + lane] = x[c];
                __syncwarp();
                const float *blkp = smem + base[it] + sd * blk[it];
                const int kp = a.dP[it] == 1 ? 4 : a.dP[it];
                float w[CH];
                pair_dense_tanh<CH>(blkp, blkp + kp * a.emb, a.dP[it], a.emb, xs, w, lane);
                if (lane < a.emb) {
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        a.out[it][o_ij + c * a.emb + lane] = w[c];
                        if (i != j) a.out[it][o_ji + c * a.emb + lane] = w[c];
                    }
                }
                if (it + 1 < a.n_iter) {
                    float y[CH];
                    const float *hwp = blkp + kp * a.emb + a.emb;
                    pair_dense_tanh<CH>(hwp, hwp + kp * a.dP[it + 1], a.dP[it], a.dP[it + 1], xs, y, lane);
                    const bool res = a.dP[it] == a.dP[it + 1];
#pragma unroll
                    for (int c = 0; c < CH; ++c) x[c] = res ? (x[c] + y[c]) * 0.70710678118654752f : y[c];
                }
            }
        }
    }
}

// Forward-only pair stream (the Metropolis step): one warp carries P unordered pairs of the SAME spin class through all iterations at
// once, so every LDS.128 of layer weights feeds 4 P FMAs instead of 4 (the single-pair version is bound by those loads).
template <int P>
__device__ __forceinline__ void pair_dense_multi(const float *__restrict__ W4, int din, int dout, const float *__restrict__ xs, float (&z)[P], int lane) {
#pragma unroll
    for (int p = 0; p < P; ++p) z[p] = 0.f;
    const bool act = lane < dout;
    if (din == 1) {
        const float w = act ? W4[lane * 4] : 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) z[p] = xs[p * 32] * w;
        return;
    }
    const int nk4 = din >> 2;
    for (int k4 = 0; k4 < nk4; ++k4) {
        const float4 w = act ? *reinterpret_cast<const float4 *>(W4 + (k4 * dout + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float4 xv = *reinterpret_cast<const float4 *>(xs + p * 32 + k4 * 4);
            z[p] = fmaf(xv.x, w.x, z[p]); z[p] = fmaf(xv.y, w.y, z[p]); z[p] = fmaf(xv.z, w.z, z[p]); z[p] = fmaf(xv.w, w.w, z[p]);
        }
    }
}

template <int P>
__global__ void __launch_bounds__(256) k_pair_stream_fwd(PairArgs a, int n_walkers) {
    extern __shared__ __align__(16) float smem[];
    int base[DPE_MAX_ITER], blk[DPE_MAX_ITER];
    int off = 0;
    for (int it = 0; it < a.n_iter; ++it) {
        base[it] = off;
        const int kp = a.dP[it] == 1 ? 4 : a.dP[it];
        blk[it] = kp * a.emb + a.emb + ((it + 1 < a.n_iter) ? kp * a.dP[it + 1] + a.dP[it + 1] : 0);
        off += 2 * blk[it];
    }
    float *xs_all = smem + off;                                        // [8 warps][P][32]
    int *pair_tab = reinterpret_cast<int *>(xs_all + 8 * P * 32);      // same-spin pairs first, then different-spin pairs; (i << 8) | j
    auto stage = [&](float *dst, const float *src, int din, int dout) {
        const int kp = din == 1 ? 4 : din;
        for (int t = threadIdx.x; t < kp * dout; t += blockDim.x) {
            int k4 = t / (dout * 4), rem = t - k4 * dout * 4, n = rem >> 2, kk = rem & 3, k = k4 * 4 + kk;
            dst[t] = k < din ? src[k * dout + n] : 0.f;
        }
    };
    for (int it = 0; it < a.n_iter; ++it)
        for (int sd = 0; sd < 2; ++sd) {
            float *dst = smem + base[it] + sd * blk[it];
            const int kp = a.dP[it] == 1 ? 4 : a.dP[it];
            stage(dst, a.ww[it][sd], a.dP[it], a.emb);
            for (int t = threadIdx.x; t < a.emb; t += blockDim.x) dst[kp * a.emb + t] = a.wb[it][sd][t];
            if (it + 1 < a.n_iter) {
                stage(dst + kp * a.emb + a.emb, a.hw[it][sd], a.dP[it], a.dP[it + 1]);
                for (int t = threadIdx.x; t < a.dP[it + 1]; t += blockDim.x) dst[kp * a.emb + a.emb + kp * a.dP[it + 1] + t] = a.hb[it][sd][t];
            }
        }
    const int N = a.N, U = a.U, D = N - U;
    const int n_same = U * (U + 1) / 2 + D * (D + 1) / 2, n_diff = U * D;
    if (threadIdx.x == 0) {                                             // tiny table, built by one thread
        int ps = 0, pd = n_same;
        for (int i = 0; i < N; ++i)
            for (int j = i; j < N; ++j) {
                if ((i < U) == (j < U)) pair_tab[ps++] = (i << 8) | j;
                else pair_tab[pd++] = (i << 8) | j;
            }
    }
    __syncthreads();
    const int gs = (n_same + P - 1) / P, gd = (n_diff + P - 1) / P, n_groups = gs + gd;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *xs = xs_all + wib * P * 32;
    const long n_units = (long)n_walkers * n_groups;
    for (long u = blockIdx.x * 8L + wib; u < n_units; u += gridDim.x * 8L) {
        const long b = u / n_groups;
        const int g = (int)(u - b * n_groups);
        const int sd = g >= gs ? 1 : 0;
        const int first = sd ? n_same + (g - gs) * P : g * P;
        const int cnt = min(P, (sd ? n_same + n_diff : n_same) - first);
        int pi[P], pj[P];
        float x[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int pk = pair_tab[first + min(p, cnt - 1)];            // padding slots repeat the last pair and are not stored
            pi[p] = pk >> 8; pj[p] = pk & 255;
            const float *ri = a.r + (b * N + pi[p]) * 3, *rj = a.r + (b * N + pj[p]) * 3;
            const float dx = rj[0] - ri[0], dy = rj[1] - ri[1], dz = rj[2] - ri[2];
            const float d = (pi[p] == pj[p]) ? 0.f : sqrtf(dx * dx + dy * dy + dz * dz);
            x[p] = lane == 0 ? d : 0.f;
        }
#pragma unroll
        for (int it = 0; it < DPE_MAX_ITER; ++it) {
            if (it < a.n_iter) {
                __syncwarp();
#pragma unroll
                for (int p = 0; p < P; ++p) xs[p * 32 + lane] = x[p];
                __syncwarp();
                const float *blkp = smem + base[it] + sd * blk[it];
                const int kp = a.dP[it] == 1 ? 4 : a.dP[it];
                float z[P];
                pair_dense_multi<P>(blkp, a.dP[it], a.emb, xs, z, lane);
                if (lane < a.emb) {
                    const float bias = blkp[kp * a.emb + lane];
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        if (p < cnt) {
                            const float w = tanh_f32(z[p] + bias);
                            a.out[it][((b * N + pi[p]) * N + pj[p]) * (long)a.emb + lane] = w;
                            if (pi[p] != pj[p]) a.out[it][((b * N + pj[p]) * N + pi[p]) * (long)a.emb + lane] = w;
                        }
                    }
                }
                if (it + 1 < a.n_iter) {
                    const float *hwp = blkp + kp * a.emb + a.emb;
                    const int dn = a.dP[it + 1];
                    pair_dense_multi<P>(hwp, a.dP[it], dn, xs, z, lane);
                    const float bias = lane < dn ? hwp[kp * dn + lane] : 0.f;
                    const bool res = a.dP[it] == dn;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float y = lane < dn ? tanh_f32(z[p] + bias) : 0.f;
                        x[p] = res ? (x[p] + y) * 0.70710678118654752f : y;
                    }
                }
            }
        }
    }
}

int launch_pair_stream(dpe_model *m, const float *r, int Bc, int CP, float *pw_base, const size_t *pw_off, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    PairArgs a;
    a.r = r; a.n_iter = d.n_iterations; a.N = d.n_el; a.U = d.n_up; a.emb = d.emb_dim;
    a.n_pairs = (long)Bc * d.n_el * d.n_el;
    size_t fl = 0;
    for (int it = 0; it < d.n_iterations; ++it) {
        const IterParams &p = m->it[it];
        a.dP[it] = p.dP;
        a.out[it] = pw_base + pw_off[it];
        a.ww[it][0] = p.w_same.w; a.wb[it][0] = p.w_same.b;
        a.ww[it][1] = p.w_diff.w; a.wb[it][1] = p.w_diff.b;
        a.hw[it][0] = p.h_same.w; a.hb[it][0] = p.h_same.b;
        a.hw[it][1] = p.h_diff.w; a.hb[it][1] = p.h_diff.b;
        const size_t kp = p.dP == 1 ? 4 : p.dP;
        fl += 2 * (kp * d.emb_dim + d.emb_dim);
        if (it + 1 < d.n_iterations) fl += 2 * (kp * m->it[it + 1].dP + m->it[it + 1].dP);
    }
    fl += (size_t)8 * CP * 32 + (size_t)d.n_el * (d.n_el + 1) / 2 + 4;
    size_t smem = fl * sizeof(float);
    long blocks = ((long)Bc * (d.n_el * (d.n_el + 1) / 2) + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    int e;
    static const bool single_pair_fwd = getenv("DPE_PAIR_FWD_SINGLE") != nullptr;     // debug: the one-pair-per-warp forward kernel
    if (CP == 1 && !single_pair_fwd) {
        // forward pass: the tensor-core pair stream (pair_tc.cu) when the layers have its 1 -> 32 -> 32 shape and the tensor-core path is on
        e = launch_pair_stream_tc(m, r, Bc, pw_base, pw_off, s);
        if (e != DPE_ERR_UNSUPPORTED) return e;
    }
    if (CP == 1 && !single_pair_fwd) {
        constexpr int P = 4;
        const int U = d.n_up, D = d.n_el - U;
        const int n_groups = (U * (U + 1) / 2 + D * (D + 1) / 2 + P - 1) / P + (U * D + P - 1) / P;
        size_t smem_f = (fl - (size_t)8 * CP * 32 + (size_t)8 * P * 32) * sizeof(float);
        long blocks_f = ((long)Bc * n_groups + 7) / 8;
        if (blocks_f > 148 * 4) blocks_f = 148 * 4;
        if ((e = opt_in_smem(m, KID_PAIR1, k_pair_stream_fwd<P>))) return e;
        k_pair_stream_fwd<P><<<(int)blocks_f, 256, smem_f, s>>>(a, Bc);
    } else if (CP == 1) {
        if ((e = opt_in_smem(m, KID_MCMC_FUSED, k_pair_stream<1>))) return e;
        k_pair_stream<1><<<(int)blocks, 256, smem, s>>>(a);
    } else {
        if ((e = opt_in_smem(m, KID_PAIR3, k_pair_stream<3>))) return e;
        k_pair_stream<3><<<(int)blocks, 256, smem, s>>>(a);
    }
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// tanh rule on a GEMM output, in place: z[group][c][0:width] (row stride ld)
//   z0 += bias + add[.,0]; y = tanh(z0); t_k = (1-y^2) z_k ; lap = (1-y^2) z_lap - 2 y (1-y^2) sum_k z_k^2
// `add` (may be null) is indexed [group / groups_per_add][c][width]  (the spin-mean addend of h_el).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_act(float *__restrict__ z, int ld, long n_groups, int C, int width,
                                              const float *__restrict__ bias, const float *__restrict__ add,
                                              int groups_per_add) {
    // one thread = 4 consecutive features of one group; channel loads are issued in batches of UB before the
    // dependent stores (the update is in place, so the compiler cannot reorder them itself)
    constexpr int UB = 8;
    const int w4 = width >> 2;
    const long total = n_groups * w4;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long g = idx / w4;
        const int f = (int)(idx - g * w4) << 2;
        float *zp = z + g * C * ld + f;
        const float *ap = add ? add + (g / groups_per_add) * C * width + f : nullptr;
        float4 z0 = *reinterpret_cast<const float4 *>(zp);
        if (bias) { float4 b = *reinterpret_cast<const float4 *>(bias + f); z0.x += b.x; z0.y += b.y; z0.z += b.z; z0.w += b.w; }
        if (ap) { float4 a = *reinterpret_cast<const float4 *>(ap); z0.x += a.x; z0.y += a.y; z0.z += a.z; z0.w += a.w; }
        float4 y = make_float4(tanh_f32(z0.x), tanh_f32(z0.y), tanh_f32(z0.z), tanh_f32(z0.w));
        *reinterpret_cast<float4 *>(zp) = y;
        if (C > 1) {
            const float4 d1 = make_float4(1.f - y.x * y.x, 1.f - y.y * y.y, 1.f - y.z * y.z, 1.f - y.w * y.w);
            float4 ssq = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = 1; c0 < C - 1; c0 += UB) {
                float4 v[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (c0 + u < C - 1) {
                        v[u] = *reinterpret_cast<const float4 *>(zp + (long)(c0 + u) * ld);
                        if (ap) {
                            float4 a = *reinterpret_cast<const float4 *>(ap + (long)(c0 + u) * width);
                            v[u].x += a.x; v[u].y += a.y; v[u].z += a.z; v[u].w += a.w;
                        }
                    }
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (c0 + u < C - 1) {
                        ssq.x = fmaf(v[u].x, v[u].x, ssq.x); ssq.y = fmaf(v[u].y, v[u].y, ssq.y);
                        ssq.z = fmaf(v[u].z, v[u].z, ssq.z); ssq.w = fmaf(v[u].w, v[u].w, ssq.w);
                        *reinterpret_cast<float4 *>(zp + (long)(c0 + u) * ld) = make_float4(d1.x * v[u].x, d1.y * v[u].y, d1.z * v[u].z, d1.w * v[u].w);
                    }
            }
            float4 zl = *reinterpret_cast<const float4 *>(zp + (long)(C - 1) * ld);
            if (ap) { float4 a = *reinterpret_cast<const float4 *>(ap + (long)(C - 1) * width); zl.x += a.x; zl.y += a.y; zl.z += a.z; zl.w += a.w; }
            *reinterpret_cast<float4 *>(zp + (long)(C - 1) * ld) =
                make_float4(d1.x * zl.x - 2.f * y.x * d1.x * ssq.x, d1.y * zl.y - 2.f * y.y * d1.y * ssq.y,
                            d1.z * zl.z - 2.f * y.z * d1.z * ssq.z, d1.w * zl.w - 2.f * y.w * d1.w * ssq.w);
        }
    }
}

// Forward pass (one channel): bias + spin-mean addend + tanh of a main-layer output, in place, one block per walker -- and, from the same registers,
// the spin means of the RESULT, which the next iteration would otherwise form with a second pass over it (k_mean).  Same summation order and
// arithmetic as k_act + k_mean: the results are bit-identical to the two-kernel sequence.
__global__ void __launch_bounds__(256) k_act_mean_fwd(float *__restrict__ z, int ld, int N, int U, int width, const float *__restrict__ bias,
                                                       const float *__restrict__ add, float *__restrict__ mean) {
    const long b = blockIdx.x;
    constexpr int CH = 16;
    for (int f = threadIdx.x; f < width; f += blockDim.x) {
        const float bia = bias ? bias[f] : 0.f, ad = add ? add[b * width + f] : 0.f;
        float up = 0.f, dn = 0.f;
        float *zp = z + b * N * (long)ld + f;
        for (int i0 = 0; i0 < N; i0 += CH) {
            float v[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k)
                if (i0 + k < N) v[k] = zp[(long)(i0 + k) * ld];
#pragma unroll
            for (int k = 0; k < CH; ++k)
                if (i0 + k < N) {
                    const float y = tanh_f32((v[k] + bia) + ad);
                    zp[(long)(i0 + k) * ld] = y;
                    if (i0 + k < U) up += y; else dn += y;
                }
        }
        if (mean) {
            mean[b * 2 * width + f] = up / (float)U;
            mean[b * 2 * width + width + f] = dn / (float)(N - U);
        }
    }
}

int launch_act_mean_fwd(dpe_model *m, float *z, int ld, int Bc, int width, const float *bias, const float *add, float *mean, cudaStream_t s) {
    k_act_mean_fwd<<<Bc, 256, 0, s>>>(z, ld, m->dims.n_el, m->dims.n_up, width, bias, add, mean);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

int launch_act(dpe_model *m, float *z, int ld, int n_groups, int C, int width, const float *bias, const float *add,
               int groups_per_add, cudaStream_t s) {
    if ((width & 3) || (ld & 3)) return set_error(DPE_ERR_UNSUPPORTED, "act: width=%d / ld=%d must be multiples of 4", width, ld);
    long total = (long)n_groups * (width >> 2);
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    k_act<<<(int)blocks, 256, 0, s>>>(z, ld, n_groups, C, width, bias, add, groups_per_add);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// spin means (ferminet_embedding.py:60-72): mean[b][c][0:d] = mean_{i<U} h, [d:2d] = mean_{i>=U} h
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mean(const float *__restrict__ x, int ldx, int Bc, int N, int U, int C, int d_in,
                                               float *__restrict__ mean) {
    const int d4 = d_in >> 2;
    const long total = (long)Bc * C * d4;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int f = (int)(idx % d4) << 2;
        const long bc = idx / d4;
        const int c = (int)(bc % C);
        const long b = bc / C;
        const float *xp = x + ((b * N) * C + c) * (long)ldx + f;
        const long stride = (long)C * ldx;
        float4 up = make_float4(0.f, 0.f, 0.f, 0.f), dn = up;
        int i = 0;
        for (; i + 4 <= U; i += 4) {
            float4 a0 = *reinterpret_cast<const float4 *>(xp + (i + 0) * stride), a1 = *reinterpret_cast<const float4 *>(xp + (i + 1) * stride);
            float4 a2 = *reinterpret_cast<const float4 *>(xp + (i + 2) * stride), a3 = *reinterpret_cast<const float4 *>(xp + (i + 3) * stride);
            up.x += a0.x; up.y += a0.y; up.z += a0.z; up.w += a0.w;
            up.x += a1.x; up.y += a1.y; up.z += a1.z; up.w += a1.w;
            up.x += a2.x; up.y += a2.y; up.z += a2.z; up.w += a2.w;
            up.x += a3.x; up.y += a3.y; up.z += a3.z; up.w += a3.w;
        }
        for (; i < U; ++i) { float4 a = *reinterpret_cast<const float4 *>(xp + i * stride); up.x += a.x; up.y += a.y; up.z += a.z; up.w += a.w; }
        for (; i + 4 <= N; i += 4) {
            float4 a0 = *reinterpret_cast<const float4 *>(xp + (i + 0) * stride), a1 = *reinterpret_cast<const float4 *>(xp + (i + 1) * stride);
            float4 a2 = *reinterpret_cast<const float4 *>(xp + (i + 2) * stride), a3 = *reinterpret_cast<const float4 *>(xp + (i + 3) * stride);
            dn.x += a0.x; dn.y += a0.y; dn.z += a0.z; dn.w += a0.w;
            dn.x += a1.x; dn.y += a1.y; dn.z += a1.z; dn.w += a1.w;
            dn.x += a2.x; dn.y += a2.y; dn.z += a2.z; dn.w += a2.w;
            dn.x += a3.x; dn.y += a3.y; dn.z += a3.z; dn.w += a3.w;
        }
        for (; i < N; ++i) { float4 a = *reinterpret_cast<const float4 *>(xp + i * stride); dn.x += a.x; dn.y += a.y; dn.z += a.z; dn.w += a.w; }
        const float nu = (float)U, nd = (float)(N - U);
        float *mp = mean + bc * 2 * d_in + f;
        *reinterpret_cast<float4 *>(mp) = make_float4(up.x / nu, up.y / nu, up.z / nu, up.w / nu);
        *reinterpret_cast<float4 *>(mp + d_in) = make_float4(dn.x / nd, dn.y / nd, dn.z / nd, dn.w / nd);
    }
}

int launch_mean(dpe_model *m, const float *x, int ldx, int Bc, int C, int d_in, float *mean, cudaStream_t s) {
    if (d_in & 3) return set_error(DPE_ERR_UNSUPPORTED, "mean: d_in=%d must be a multiple of 4", d_in);
    long total = (long)Bc * C * (d_in >> 2);
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    k_mean<<<(int)blocks, 256, 0, s>>>(x, ldx, Bc, m->dims.n_el, m->dims.n_up, C, d_in, mean);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// SchNet convolution (ferminet_embedding.py:159-175) with the product rule:
//   conv_ee[i] = sum_j w[i,j] * hm[j]   (incl. j == i),   conv_eI[i] = (precomputed by the el-ion stream)
// Laplacian mode: one block per (walker, slab of CS channels); thread = (channel, feature). The channel slab of
// hm for all electrons stays in shared memory; the pair weights of one electron i are staged per iteration.
// ------------------------------------------------------------------------------------------------
// Dense part: conv_ee[b,i,c,f] = sum_j w[b,i,j,f] * hm[b,j,c,f] is, per (walker, feature), an [N x N] x [N x C] product.
// Block = (walker, 8 features, 32 channels); w and the hm slab are staged once in shared memory; each thread owns one
// (channel, feature) and register-blocks 4 electrons i (one LDS.128 of w + one LDS of hm per 4 FMAs).
constexpr int CV_FG = 8, CV_CB = 32;
__global__ void __launch_bounds__(CV_FG * CV_CB) k_conv_dense(int N, int C, int CP, int emb, const float *__restrict__ hm,
                                                              const float *__restrict__ pw, float *__restrict__ x, int ldx,
                                                              int col_ee) {
    extern __shared__ __align__(16) float sm[];
    int NP = (N + 3) & ~3;
    if ((NP & 7) == 0) NP += 4;                   // row stride = 4 (mod 8) floats: the 8 features of a quarter warp hit distinct banks
    float *A_s = sm;                              // [j][f][NP]   w(i, j, f), i fastest
    float *B_s = A_s + N * CV_FG * NP;            // [j][cl][f]
    const int n_fg = emb / CV_FG, n_cb = (C + CV_CB - 1) / CV_CB;
    int bid = blockIdx.x;
    const int cb = bid % n_cb; bid /= n_cb;
    const int fg = bid % n_fg;
    const long b = bid / n_fg;
    const int f0 = fg * CV_FG, c0 = cb * CV_CB;
    const int tid = threadIdx.x, f = tid & (CV_FG - 1), cl = tid >> 3;
    {   // 32 threads x 8 features walk j for every i: no integer division in the staging loops
        const int ff = tid & (CV_FG - 1), jl = tid >> 3;
        for (int i = 0; i < N; ++i) {
            const float *src = pw + ((b * N + i) * N) * (long)CP * emb + f0 + ff;
            for (int j = jl; j < N; j += CV_CB) A_s[(j * CV_FG + ff) * NP + i] = src[(long)j * CP * emb];
        }
    }
    if (NP > N)
        for (int t = tid; t < N * CV_FG * (NP - N); t += blockDim.x) {
            const int jf = t / (NP - N), i = N + (t - jf * (NP - N));
            A_s[jf * NP + i] = 0.f;
        }
    for (int j = 0; j < N; ++j)      // thread = (channel cl, feature f) of the slab
        B_s[(j * CV_CB + cl) * CV_FG + f] = (c0 + cl < C) ? hm[((b * N + j) * (long)C + c0 + cl) * emb + f0 + f] : 0.f;
    __syncthreads();
    const int c = c0 + cl;
    if (c >= C) return;
    const float *ap = A_s + f * NP;
    const float *bp = B_s + cl * CV_FG + f;
    float *xp = x + ((b * N) * (long)C + c) * ldx + col_ee + f0 + f;
    const long xstride = (long)C * ldx;
    for (int ig = 0; ig < N; ig += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int j = 0; j < N; ++j) {
            const float4 w = *reinterpret_cast<const float4 *>(ap + j * CV_FG * NP + ig);
            const float h = bp[j * CV_CB * CV_FG];
            a0 = fmaf(w.x, h, a0); a1 = fmaf(w.y, h, a1); a2 = fmaf(w.z, h, a2); a3 = fmaf(w.w, h, a3);
        }
        xp[(ig + 0) * xstride] = a0;
        if (ig + 1 < N) xp[(ig + 1) * xstride] = a1;
        if (ig + 2 < N) xp[(ig + 2) * xstride] = a2;
        if (ig + 3 < N) xp[(ig + 3) * xstride] = a3;
    }
}

// Second version of the dense part.  The first one (k_conv_dense) walks j once per block of four electrons i and stages its
// operands twice (two channel blocks): ncu counted 11 instructions per useful FMA.  Here a block = (walker, 8 features) covers ALL
// channels, a thread = (channel, feature) keeps up to 16 electrons i in registers and walks j once: per j one LDS of hm and four
// LDS.128 of w feed 16 FMAs.  Shared memory: w as [j][f][NP] (i fastest, NP = 20: the 8 features of a quarter warp hit distinct
// banks), the hm slab as [j][c][f].
constexpr int CV2_NP = 20;
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_conv_dense2(int N, int C, int CP, int emb, const float *__restrict__ hm,
                                                      const float *__restrict__ pw, float *__restrict__ x, int ldx, int col_ee) {
    extern __shared__ __align__(16) float sm[];
    float *A_s = sm;                               // [j][f][NP]  (one block of <= 16 electrons i at a time)
    float *B_s = A_s + N * CV_FG * CV2_NP;         // [j][c][f]
    const int n_fg = emb / CV_FG;
    const int fg = blockIdx.x % n_fg;
    const long b = blockIdx.x / n_fg;
    const int f0 = fg * CV_FG;
    const int tid = threadIdx.x, f = tid & (CV_FG - 1), c = tid >> 3;       // blockDim.x = 8 * C
    static_assert(CV_FG == 8, "the float4 staging below assumes 8 features per block");
    for (int t = tid; t < N * C * 2; t += blockDim.x) {                   // 16-byte staging loads (two float4 per row of 8 features)
        const int row = t >> 1, hq = (t & 1) * 4;                           // row = j * C + c
        *reinterpret_cast<float4 *>(B_s + row * CV_FG + hq) = *reinterpret_cast<const float4 *>(hm + (b * N * C + row) * (long)emb + f0 + hq);
    }
    for (int i0 = 0; i0 < N; i0 += 16) {
        const int ni = min(16, N - i0);
        __syncthreads();                           // previous i block consumed (and, first time, nothing)
        for (int t = tid; t < N * CV2_NP * 2; t += blockDim.x) {           // one float4 of w per thread, transposed into A_s[j][f][i]
            const int hq = (t & 1) * 4, i = (t >> 1) % CV2_NP, j = (t >> 1) / CV2_NP;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < ni) v = *reinterpret_cast<const float4 *>(pw + (((b * N + i0 + i) * N) + j) * (long)CP * emb + f0 + hq);
            float *dst = A_s + (j * CV_FG + hq) * CV2_NP + i;
            dst[0] = v.x; dst[CV2_NP] = v.y; dst[2 * CV2_NP] = v.z; dst[3 * CV2_NP] = v.w;
        }
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        const float *ap = A_s + f * CV2_NP;
        const float *bp = B_s + c * CV_FG + f;
        for (int j = 0; j < N; ++j) {
            const float h = bp[j * C * CV_FG];
            const float4 w0 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP);
            const float4 w1 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP + 4);
            const float4 w2 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP + 8);
            const float4 w3 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP + 12);
            acc[0] = fmaf(w0.x, h, acc[0]); acc[1] = fmaf(w0.y, h, acc[1]); acc[2] = fmaf(w0.z, h, acc[2]); acc[3] = fmaf(w0.w, h, acc[3]);
            acc[4] = fmaf(w1.x, h, acc[4]); acc[5] = fmaf(w1.y, h, acc[5]); acc[6] = fmaf(w1.z, h, acc[6]); acc[7] = fmaf(w1.w, h, acc[7]);
            acc[8] = fmaf(w2.x, h, acc[8]); acc[9] = fmaf(w2.y, h, acc[9]); acc[10] = fmaf(w2.z, h, acc[10]); acc[11] = fmaf(w2.w, h, acc[11]);
            acc[12] = fmaf(w3.x, h, acc[12]); acc[13] = fmaf(w3.y, h, acc[13]); acc[14] = fmaf(w3.z, h, acc[14]); acc[15] = fmaf(w3.w, h, acc[15]);
        }
        float *xp = x + (((b * N + i0) * (long)C) + c) * ldx + col_ee + f0 + f;
        const long xstride = (long)C * ldx;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < ni) xp[i * xstride] = acc[i];
    }
}

// Third version: the sparse product-rule terms (what k_conv_special adds in a second read-modify-write pass) are folded into the dense
// kernel, so conv_ee / conv_eI are written exactly once.  Block = (walker, 8 features), thread = (channel c, feature f) as in
// k_conv_dense2.  A prologue run by the first N x 8 threads computes, per electron i and feature, the sums over j that only the
// own-electron tangent channels and the Laplacian channel need,
//     s_a[i]  = sum_{j != i} w'_ij u_ij,a hm0_j
//     lap[i]  = sum_{j != i} (2 w''_ij + 4 w'_ij / d_ij) hm0_j + 2 w'_ij u_ij . (d hm_j / d r_j - d hm_j / d r_i)
// from the hm slab that is resident in shared memory anyway; every other tangent channel (electron e != i, axis a) gets its single
// term w'_ie u_ie,a hm0_e in the thread that owns the channel.  The thread also expands the 5-channel el-ion convolution into its
// channel of conv_eI.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_conv_fused(int N, int C, int emb, int dE, const float *__restrict__ r, const float *__restrict__ hm,
                                                     const float *__restrict__ pw, const float *__restrict__ ei, float *__restrict__ x, int ldx,
                                                     int col_ee) {
    extern __shared__ __align__(16) float sm[];
    float *A_s = sm;                               // [j][f][NP]  (one block of <= 16 electrons i at a time)
    float *B_s = A_s + N * CV_FG * CV2_NP;         // [j][c][f]
    float *W1_s = B_s + N * C * CV_FG;             // [i][j][f]  w'
    float *W2_s = W1_s + N * N * CV_FG;            // [i][j][f]  w''
    float4 *U_s = reinterpret_cast<float4 *>(W2_s + N * N * CV_FG);    // [i][j] (u_x, u_y, u_z, 1/d) of r_j - r_i
    float *S_s = reinterpret_cast<float *>(U_s + N * N);               // [i][a][f]
    float *L_s = S_s + N * 3 * CV_FG;              // [i][f]
    const int n_fg = emb / CV_FG;
    const int fg = blockIdx.x % n_fg;
    const long b = blockIdx.x / n_fg;
    const int f0 = fg * CV_FG;
    const int tid = threadIdx.x, f = tid & (CV_FG - 1), c = tid >> 3;       // blockDim.x = 8 * C
    // global -> shared with 16-byte accesses (the 8 features of a block are two float4 per row): a third of the load instructions of the scalar version
    static_assert(CV_FG == 8, "the float4 staging below assumes 8 features per block");
    for (int t = tid; t < N * C * 2; t += blockDim.x) {
        const int row = t >> 1, hq = (t & 1) * 4;                           // row = j * C + c
        *reinterpret_cast<float4 *>(B_s + row * CV_FG + hq) = *reinterpret_cast<const float4 *>(hm + (b * N * C + row) * (long)emb + f0 + hq);
    }
    for (int t = tid; t < N * N * 2; t += blockDim.x) {
        const int ij = t >> 1, hq = (t & 1) * 4;
        const float *p1 = pw + (b * N * N + ij) * 3L * emb + emb + f0 + hq;
        *reinterpret_cast<float4 *>(W1_s + ij * CV_FG + hq) = *reinterpret_cast<const float4 *>(p1);
        *reinterpret_cast<float4 *>(W2_s + ij * CV_FG + hq) = *reinterpret_cast<const float4 *>(p1 + emb);
    }
    for (int t = tid; t < N * N; t += blockDim.x) {
        const int i = t / N, j = t - i * N;
        const float *rb = r + b * N * 3;
        const float dx = rb[3 * j] - rb[3 * i], dy = rb[3 * j + 1] - rb[3 * i + 1], dz = rb[3 * j + 2] - rb[3 * i + 2];
        const float inv = i == j ? 0.f : 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        U_s[t] = make_float4(dx * inv, dy * inv, dz * inv, inv);
    }
    // the w slab of the first block of electrons does not depend on the prologue: fetch it in the same global-load phase
    auto load_w = [&](int i0, int ni) {
        for (int t = tid; t < N * CV2_NP * 2; t += blockDim.x) {           // one float4 of w per thread, transposed into A_s[j][f][i]
            const int hq = (t & 1) * 4, i = (t >> 1) % CV2_NP, j = (t >> 1) / CV2_NP;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < ni) v = *reinterpret_cast<const float4 *>(pw + (((b * N + i0 + i) * N) + j) * 3L * emb + f0 + hq);
            float *dst = A_s + (j * CV_FG + hq) * CV2_NP + i;
            dst[0] = v.x; dst[CV2_NP] = v.y; dst[2 * CV2_NP] = v.z; dst[3 * CV2_NP] = v.w;
        }
    };
    load_w(0, min(16, N));
    __syncthreads();
    if (tid < N * CV_FG) {                          // prologue: thread = (electron i, feature f)
        const int i = tid >> 3;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, lap = 0.f;
        for (int j = 0; j < N; ++j) {
            const float4 u = U_s[i * N + j];         // all zero for j == i: every term below vanishes
            const float w1 = W1_s[(i * N + j) * CV_FG + f], w2 = j == i ? 0.f : W2_s[(i * N + j) * CV_FG + f];
            const float *hj = B_s + (j * C) * CV_FG + f;
            const float h0 = hj[0];
            const float *hjj = hj + (1 + 3 * j) * CV_FG, *hji = hj + (1 + 3 * i) * CV_FG;
            const float cross = u.x * (hjj[0] - hji[0]) + u.y * (hjj[CV_FG] - hji[CV_FG]) + u.z * (hjj[2 * CV_FG] - hji[2 * CV_FG]);
            const float t = w1 * h0;
            s0 = fmaf(t, u.x, s0); s1 = fmaf(t, u.y, s1); s2 = fmaf(t, u.z, s2);
            lap += (2.f * w2 + 4.f * w1 * u.w) * h0 + 2.f * w1 * cross;
        }
        S_s[(i * 3 + 0) * CV_FG + f] = s0; S_s[(i * 3 + 1) * CV_FG + f] = s1; S_s[(i * 3 + 2) * CV_FG + f] = s2;
        L_s[i * CV_FG + f] = lap;
    }
    const bool tangent = c >= 1 && c < C - 1;
    const int e = tangent ? (c - 1) / 3 : 0, ax = tangent ? (c - 1) - 3 * e : 0;
    const float h0e = B_s[(e * C) * CV_FG + f];
    for (int i0 = 0; i0 < N; i0 += 16) {
        const int ni = min(16, N - i0);
        __syncthreads();                           // previous i block consumed; first time: prologue tables complete
        if (i0 > 0) {
            load_w(i0, ni);
            __syncthreads();
        }
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        const float *ap = A_s + f * CV2_NP;
        const float *bp = B_s + c * CV_FG + f;
        for (int j = 0; j < N; ++j) {
            const float h = bp[j * C * CV_FG];
            const float4 w0 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP);
            const float4 w1 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP + 4);
            const float4 w2 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP + 8);
            const float4 w3 = *reinterpret_cast<const float4 *>(ap + j * CV_FG * CV2_NP + 12);
            acc[0] = fmaf(w0.x, h, acc[0]); acc[1] = fmaf(w0.y, h, acc[1]); acc[2] = fmaf(w0.z, h, acc[2]); acc[3] = fmaf(w0.w, h, acc[3]);
            acc[4] = fmaf(w1.x, h, acc[4]); acc[5] = fmaf(w1.y, h, acc[5]); acc[6] = fmaf(w1.z, h, acc[6]); acc[7] = fmaf(w1.w, h, acc[7]);
            acc[8] = fmaf(w2.x, h, acc[8]); acc[9] = fmaf(w2.y, h, acc[9]); acc[10] = fmaf(w2.z, h, acc[10]); acc[11] = fmaf(w2.w, h, acc[11]);
            acc[12] = fmaf(w3.x, h, acc[12]); acc[13] = fmaf(w3.y, h, acc[13]); acc[14] = fmaf(w3.z, h, acc[14]); acc[15] = fmaf(w3.w, h, acc[15]);
        }
        float *xp = x + (((b * N + i0) * (long)C) + c) * ldx + col_ee + f0 + f;
        const long xstride = (long)C * ldx;
        const bool do_ei = f0 + f < dE;
        const float *eip = ei + ((b * N + i0) * 5L) * dE + f0 + f;                        // [i][5][dE]
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < ni) {
                const int ig = i0 + i;
                float v = acc[i];
                if (tangent) {
                    if (e == ig) {
                        v -= S_s[(ig * 3 + ax) * CV_FG + f];
                    } else {
                        const float4 u = U_s[ig * N + e];
                        v += (W1_s[(ig * N + e) * CV_FG + f] * h0e) * (ax == 0 ? u.x : (ax == 1 ? u.y : u.z));
                    }
                } else if (c == C - 1) {
                    v += L_s[ig * CV_FG + f];
                }
                xp[i * xstride] = v;
                if (do_ei) {        // conv_eI: the 5-channel el-ion convolution expanded into the C channels of electron i
                    float w = 0.f;
                    if (c == 0) w = eip[(long)i * 5 * dE];
                    else if (c == C - 1) w = eip[(long)i * 5 * dE + 4 * dE];
                    else if (e == ig) w = eip[(long)i * 5 * dE + (1 + ax) * dE];
                    xp[i * xstride + emb] = w;
                }
            }
    }
}

// Sparse part of the product rule (w depends on r_i, r_j only) and the expansion of conv_eI: one thread per (walker, i, f).
//   d/dr_j  (j != i):  + w'_ij u_ij hm_j          d/dr_i:  - sum_j w'_ij u_ij hm_j
//   Laplacian:  sum_j [ (2 w''_ij + 4 w'_ij / d_ij) hm_j + 2 w'_ij u_ij . (d hm_j/d r_j - d hm_j/d r_i) ]
__global__ void __launch_bounds__(256, 6) k_conv_special(const float *__restrict__ r, long n_rows, int N, int C, int emb, int dE,
                                                       const float *__restrict__ hm, const float *__restrict__ pw,
                                                       const float *__restrict__ ei, float *__restrict__ x, int ldx, int col_ee) {
    const long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);     // (b, i)
    const int f = threadIdx.x & 31;
    if (row >= n_rows) return;
    const long b = row / N;
    const int i = (int)(row - b * N);
    const float *rb = r + b * N * 3;
    const float xi = rb[3 * i], yi = rb[3 * i + 1], zi = rb[3 * i + 2];
    float *xr = x + row * (long)C * ldx + col_ee;            // channel c at xr[c * ldx]
    if (f < emb) {
        const float *pwi = pw + row * (long)N * 3 * emb + f;  // [j][ch][f]
        const float *hmb = hm + b * N * (long)C * emb + f;    // [j][c][f]
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, lap = 0.f;
        for (int j = 0; j < N; ++j) {
            if (j == i) continue;
            const float dx = rb[3 * j] - xi, dy = rb[3 * j + 1] - yi, dz = rb[3 * j + 2] - zi;
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            const float ux = dx * inv, uy = dy * inv, uz = dz * inv;
            const float w1 = pwi[(long)(j * 3 + 1) * emb], w2 = pwi[(long)(j * 3 + 2) * emb];
            const float *hj = hmb + (long)j * C * emb;
            const float h0 = hj[0];
            const float *hjj = hj + (long)(1 + 3 * j) * emb, *hji = hj + (long)(1 + 3 * i) * emb;
            const float cross = ux * (hjj[0] - hji[0]) + uy * (hjj[emb] - hji[emb]) + uz * (hjj[2 * emb] - hji[2 * emb]);
            const float t = w1 * h0;
            float *xt = xr + (long)(1 + 3 * j) * ldx + f;
            xt[0] += t * ux; xt[ldx] += t * uy; xt[2 * ldx] += t * uz;
            s0 = fmaf(t, ux, s0); s1 = fmaf(t, uy, s1); s2 = fmaf(t, uz, s2);
            lap += (2.f * w2 + 4.f * w1 * inv) * h0 + 2.f * w1 * cross;
        }
        float *xs = xr + (long)(1 + 3 * i) * ldx + f;
        xs[0] -= s0; xs[ldx] -= s1; xs[2 * ldx] -= s2;
        xr[(long)(C - 1) * ldx + f] += lap;
    }
    if (f < dE) {          // conv_eI: the 5-channel el-ion convolution expanded into the C channels of electron i
        const float *eii = ei + row * 5 * dE + f;
        float *xe = xr + emb + f;
        for (int c = 0; c < C; ++c) {
            float v = 0.f;
            if (c == 0) v = eii[0];
            else if (c == C - 1) v = eii[4 * dE];
            else if (c >= 1 + 3 * i && c < 4 + 3 * i) v = eii[(c - 3 * i) * dE];
            xe[(long)c * ldx] = v;
        }
    }
}

// forward-only mode: one block per walker, warp per electron
__global__ void __launch_bounds__(256) k_conv_fwd(int N, int emb, int dE, const float *__restrict__ hm,
                                                   const float *__restrict__ pw, const float *__restrict__ ei,
                                                   float *__restrict__ x, int ldx, int col_ee) {
    extern __shared__ float sm[];            // hm[N][emb]
    const long b = blockIdx.x;
    for (int t = threadIdx.x; t < N * emb; t += blockDim.x) sm[t] = hm[b * N * emb + t];
    __syncthreads();
    const int f = threadIdx.x & 31;
    for (int i = threadIdx.x >> 5; i < N; i += blockDim.x >> 5) {
        float *xrow = x + (b * N + i) * (long)ldx + col_ee;
        if (f < emb) {
            const float *pwi = pw + (b * N + i) * (long)N * emb + f;
            float acc = 0.f;
            for (int j = 0; j < N; ++j) acc = fmaf(pwi[j * emb], sm[j * emb + f], acc);
            xrow[f] = acc;
        }
        if (f < dE) xrow[emb + f] = ei[(b * N + i) * dE + f];
    }
}

int launch_conv(dpe_model *m, int it, const float *r, int Bc, int C, const float *hm, const float *pw, const float *ei,
                float *x, int ldx, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const IterParams &p = m->it[it];
    const int N = d.n_el, emb = d.emb_dim;
    if (C == 1) {
        k_conv_fwd<<<Bc, 256, (size_t)N * emb * sizeof(float), s>>>(N, emb, p.dE, hm, pw, ei, x, ldx, p.d_in);
    } else {
        if (emb % CV_FG) return set_error(DPE_ERR_UNSUPPORTED, "conv: emb_dim=%d must be a multiple of %d", emb, CV_FG);
        static const bool old_dense = getenv("DPE_CONV_V1") != nullptr;
        static const bool split_conv = getenv("DPE_CONV_SPLIT") != nullptr;       // debug: dense + sparse kernels instead of the fused one
        const size_t smem3 = ((size_t)N * CV_FG * CV2_NP + (size_t)N * C * CV_FG + (size_t)2 * N * N * CV_FG + (size_t)4 * N * N + (size_t)N * 4 * CV_FG) * sizeof(float);
        // conv_eI has dE <= emb features (validate_dims), so the emb / 8 feature groups of a walker cover it
        if (!old_dense && !split_conv && C * CV_FG <= 1024 && smem3 <= (size_t)DPE_SMEM_OPTIN - 1024 && p.dE <= emb) {
            static const int minb = getenv("DPE_CONV_MINB") ? atoi(getenv("DPE_CONV_MINB")) : 4;
            if (C * CV_FG <= 384 && smem3 <= 48 * 1024) {
                if (minb == 3) k_conv_fused<384, 3><<<Bc * (emb / CV_FG), C * CV_FG, smem3, s>>>(N, C, emb, p.dE, r, hm, pw, ei, x, ldx, p.d_in);
                else k_conv_fused<384, 4><<<Bc * (emb / CV_FG), C * CV_FG, smem3, s>>>(N, C, emb, p.dE, r, hm, pw, ei, x, ldx, p.d_in);
            } else {
                if (int e = opt_in_smem(m, KID_GRAD_A, k_conv_fused<1024, 1>)) return e;
                k_conv_fused<1024, 1><<<Bc * (emb / CV_FG), C * CV_FG, smem3, s>>>(N, C, emb, p.dE, r, hm, pw, ei, x, ldx, p.d_in);
            }
            DPE_LAUNCH_CHECK(m);
            return DPE_OK;
        }
        const size_t smem2 = ((size_t)N * CV_FG * CV2_NP + (size_t)N * C * CV_FG) * sizeof(float);
        if (!old_dense && C * CV_FG <= 1024 && smem2 <= 200 * 1024) {
            // small systems: 384-thread blocks at 3 blocks per SM (the kernel is latency bound: occupancy matters more than registers)
            if (C * CV_FG <= 384 && smem2 <= 48 * 1024) {
                k_conv_dense2<384, 4><<<Bc * (emb / CV_FG), C * CV_FG, smem2, s>>>(N, C, 3, emb, hm, pw, x, ldx, p.d_in);
            } else {
                if (int e = opt_in_smem(m, KID_CONV2_BIG, k_conv_dense2<1024, 1>)) return e;
                k_conv_dense2<1024, 1><<<Bc * (emb / CV_FG), C * CV_FG, smem2, s>>>(N, C, 3, emb, hm, pw, x, ldx, p.d_in);
            }
        } else {
            int NP = (N + 3) & ~3;
            if ((NP & 7) == 0) NP += 4;
            size_t smem = ((size_t)N * CV_FG * NP + (size_t)N * CV_CB * CV_FG) * sizeof(float);
            if (smem > DPE_SMEM_OPTIN) return set_error(DPE_ERR_UNSUPPORTED, "conv: %zu bytes of shared memory", smem);
            if (int e = opt_in_smem(m, KID_CONV1, k_conv_dense)) return e;
            int n_cb = (C + CV_CB - 1) / CV_CB;
            k_conv_dense<<<Bc * (emb / CV_FG) * n_cb, CV_FG * CV_CB, smem, s>>>(N, C, 3, emb, hm, pw, x, ldx, p.d_in);
        }
        DPE_LAUNCH_CHECK(m);
        long n_rows = (long)Bc * N;
        k_conv_special<<<(int)((n_rows + 7) / 8), 256, 0, s>>>(r, n_rows, N, C, emb, p.dE, hm, pw, ei, x, ldx, p.d_in);
    }
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------
// parameter / geometry preparation
// ------------------------------------------------------------------------------------------------
__global__ void k_softplus(const float *__restrict__ x, float *__restrict__ y, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) {
        float v = x[i];
        y[i] = fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));   // jax.nn.softplus = logaddexp(x, 0)
    }
}

int launch_prepare_params(dpe_model *m, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    for (int it = 0; it < d.n_iterations; ++it) {
        IterParams &p = m->it[it];
        const float *w = p.h_el.w;   // rows: h_one (d_in) | mean_up (d_in) | mean_dn (d_in) | conv_ee (emb) | conv_eI (dE)
        size_t row = (size_t)p.d_out * sizeof(float);
        DPE_CUDA(cudaMemcpyAsync(p.w_main, w, p.d_in * row, cudaMemcpyDeviceToDevice, s));
        DPE_CUDA(cudaMemcpyAsync(p.w_main + (size_t)p.d_in * p.d_out, w + (size_t)3 * p.d_in * p.d_out,
                                 (d.emb_dim + p.dE) * row, cudaMemcpyDeviceToDevice, s));
        DPE_CUDA(cudaMemcpyAsync(p.w_mean, w + (size_t)p.d_in * p.d_out, 2 * p.d_in * row, cudaMemcpyDeviceToDevice, s));
    }
    long n = (long)d.n_ion * d.n_dets * d.n_el;
    for (int sp = 0; sp < 2 && !d.use_taos; ++sp) {
        k_softplus<<<(int)((n + 255) / 256), 256, 0, s>>>(m->alpha[sp], m->sp_alpha[sp], n);
        DPE_LAUNCH_CHECK(m);
    }
    return DPE_OK;
}

// him^it[J][f] = tanh(h_ion[Z_J - z_min] @ W + b)   (ferminet_embedding.py:171-174, input_features.py:184-191)
__global__ void k_him(const float *__restrict__ emb_tab, const float *__restrict__ Zf, int z_min, int F, int I, int dE,
                      const float *__restrict__ W, const float *__restrict__ b, float *__restrict__ him) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= I * dE) return;
    int J = idx / dE, f = idx - J * dE;
    const float *h = emb_tab + (long)((int)Zf[J] - z_min) * F;
    float acc = 0.f;
    for (int k = 0; k < F; ++k) acc = fmaf(h[k], W[k * dE + f], acc);
    him[idx] = tanh_f32(acc + b[f]);
}

// Device-side set_geometry: R / Z already live on the GPU (MCMCState.R, .Z); copies them, converts Z to float (clamped to the
// embedding vocabulary, a flag records a violation) and sums the ion-ion repulsion in the host routine's order (same float32 bits).
__global__ void k_geometry_from_device(const float *__restrict__ R, const int32_t *__restrict__ Z, int I, int z_min, int z_max,
                                       float *__restrict__ R_out, float *__restrict__ Zf, float *__restrict__ eii, int32_t *__restrict__ flag) {
    for (int t = threadIdx.x; t < 3 * I; t += blockDim.x) R_out[t] = R[t];
    for (int t = threadIdx.x; t < I; t += blockDim.x) {
        int z = Z[t];
        if (z < z_min || z > z_max) { atomicExch(flag, 1); z = min(max(z, z_min), z_max); }
        Zf[t] = (float)z;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float e = 0.f;
        for (int a = 0; a < I; ++a)
            for (int b = a + 1; b < I; ++b) {
                const float dx = R[a * 3] - R[b * 3], dy = R[a * 3 + 1] - R[b * 3 + 1], dz = R[a * 3 + 2] - R[b * 3 + 2];
                e = __fadd_rn(e, __fdiv_rn(__fmul_rn(Zf[a], Zf[b]), __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)))));
            }
        eii[0] = e;
    }
}

int launch_geometry_from_device(dpe_model *m, const float *R_dev, const int32_t *Z_dev, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    k_geometry_from_device<<<1, 128, 0, s>>>(R_dev, Z_dev, d.n_ion, d.z_min, d.z_max, m->R_dev, m->Z_dev, m->eii_dev, m->geom_flag_dev);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

int launch_prepare_geometry(dpe_model *m, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    for (int it = 0; it < d.n_iterations; ++it) {
        const IterParams &p = m->it[it];
        int n = d.n_ion * p.dE;
        k_him<<<(n + 127) / 128, 128, 0, s>>>(m->h_ion_emb, m->Z_dev, d.z_min, d.n_ion_features, d.n_ion, p.dE,
                                              p.h_ion_map.w, p.h_ion_map.b, p.him);
        DPE_LAUNCH_CHECK(m);
    }
    return DPE_OK;
}

}  // namespace dpe
