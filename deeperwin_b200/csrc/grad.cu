// Parameter gradient of the VMC loss and KFAC statistics: the backward pass of log psi^2.
//
// Reference: optimization/loss_function.py:112-154 -- the gradient of total_energy is the backward pass of
//     sum_b c_b log psi^2_b,   c_b = (E_clipped_b - mean E_clipped) / B                       (custom jvp, :143-154)
// and the KFAC curvature statistics of every dense layer (custom_kfac_jax/kfac_jax/_src/curvature_blocks.py:1594-1624 with the
// repeated-dense folding of curvature_tags_and_blocks.py:41-64; loss registered on 1/2 log psi^2 with variance 1/2, fisher_exact):
//     A = [x, 1]^T [x, 1] / B',   G = dy^T dy / B',   dy = (1 / sqrt 2) d log psi^2 / dy per sample,   B' = rows of x.
// The backward pass is linear in the per-walker cotangent, so ONE pass with unit cotangent serves both: it materialises, per
// dense layer, the layer inputs x and the output cotangents dy (one row per walker x electron / ordered pair / electron-ion pair),
// and three products of the same kernel finish the job:
//     dW = [x, 1]^T diag(c_walker(row)) dy,      A = [x, 1]^T [x, 1],      G = dy^T dy / 2.
// Everything here runs on the value channel only (C = 1): the FLOPs are 1 / (3N + 2) of the forward-Laplacian pass.
#include <cstdio>
#include <cstring>
#include "dpe_internal.cuh"

namespace dpe {

static inline size_t gr_align(size_t x) { return (x + 255) / 256 * 256; }

// ------------------------------------------------------------------------------------------------ generic products
// C[r, k] (+)= sum_n A[r, n] * W[k, n]            (backward of y = x W: dx = dy W^T with W [K, N] row-major, no transposed copy)
// Rows may be segmented (the spin block of every walker): logical row m -> (m / seg_len) * seg_stride + seg_off + m % seg_len.
struct RowMap { int seg_len, seg_stride, seg_off; };
__device__ __forceinline__ long map_row(const RowMap &rm, long m) { return rm.seg_len ? (m / rm.seg_len) * rm.seg_stride + rm.seg_off + m % rm.seg_len : m; }

__global__ void __launch_bounds__(256) k_gemm_nt(const float *__restrict__ A, long lda, const float *__restrict__ W, long ldw, float *__restrict__ C, long ldc,
                                                 int R, int K, int Nn, int accumulate, RowMap rm) {
    __shared__ float As[16][64 + 4], Ws[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long r0 = blockIdx.y * 64L;
    const int k0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int n0 = 0; n0 < Nn; n0 += 16) {
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            const int row = t >> 4, n = t & 15;
            As[n][row] = (r0 + row < R && n0 + n < Nn) ? A[map_row(rm, r0 + row) * lda + n0 + n] : 0.f;
            Ws[n][row] = (k0 + row < K && n0 + n < Nn) ? W[(long)(k0 + row) * ldw + n0 + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            float a[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[n][ty * 4 + u]; w[u] = Ws[n][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], w[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const long r = r0 + ty * 4 + u;
            const int k = k0 + tx * 4 + v;
            if (r < R && k < K) {
                float *c = C + map_row(rm, r) * ldc + k;
                *c = (accumulate ? *c : 0.f) + acc[u][v];
            }
        }
}

static int gemm_nt(dpe_model *m, const float *A, long lda, const float *W, long ldw, float *C, long ldc, long R, int K, int Nn, bool acc, cudaStream_t s,
                   RowMap rm = RowMap{0, 0, 0}) {
    if (R <= 0 || K <= 0 || Nn <= 0) return DPE_OK;
    if (!acc && ldw == Nn && R >= 1024 && R < (1L << 31) && lda < (1L << 31) && ldc < (1L << 31)) {
        // the layer's own weight read transposed, on the tensor cores (registered with tr in dpe_model_create); other shapes: FP32 cores
        const int e = dense_gemm_t(m, A, (int)lda, W, C, (int)ldc, (int)R, K, Nn, rm.seg_len, rm.seg_stride, rm.seg_off, s);
        if (e != DPE_ERR_UNSUPPORTED) return e;
    }
    dim3 grid((K + 63) / 64, (unsigned)((R + 63) / 64));
    k_gemm_nt<<<grid, 256, 0, s>>>(A, lda, W, ldw, C, ldc, (int)R, K, Nn, acc ? 1 : 0, rm);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

// P[split][m, n] = sum_{r in split} wrow(r) * Aaug[r, m] * B[r, n],    Aaug = [A (Ma columns) | 1 (if ones)]
// rows may be filtered by the spin class of an ordered electron pair (row % (N N) = i N + j; sel 0: same spin, 1: different, -1: all);
// wrow(r) = wts[r / rpw] (the per-walker cotangent) or 1.
struct AtbArgs {
    const float *A; long lda; int Ma; int ones;
    const float *B; long ldb; int Nb;
    long rows; int rpw; const float *wts;
    int sel, N, U;
    int n_split; long rows_per_split;
    float *part;       // [n_split][Ma + ones][Nb]
    RowMap rm;         // physical row of A and B for logical row r
};

__global__ void __launch_bounds__(256) k_atb(AtbArgs a) {
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int Mt = a.Ma + a.ones;
    const int n_nb = (a.Nb + 63) / 64;
    const int m0 = (blockIdx.x / n_nb) * 64, n0 = (blockIdx.x % n_nb) * 64;
    const long r_begin = blockIdx.y * a.rows_per_split, r_end = min(a.rows, r_begin + a.rows_per_split);
    float acc[4][4] = {};
    __shared__ float wrow[16];
    __shared__ long prow[16];
    for (long rc = r_begin; rc < r_end; rc += 16) {
        if (threadIdx.x < 16) {                     // per row: spin-class filter, walker weight, physical row
            const long r = rc + threadIdx.x;
            bool ok = r < r_end;
            if (ok && a.sel >= 0) {
                const int pidx = (int)(r % ((long)a.N * a.N)), i = pidx / a.N, j = pidx - i * a.N;
                ok = (((i < a.U) == (j < a.U)) ? 0 : 1) == a.sel;
            }
            wrow[threadIdx.x] = ok ? (a.wts ? a.wts[r / a.rpw] : 1.f) : 0.f;
            prow[threadIdx.x] = ok ? map_row(a.rm, r) : -1;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            const int rr = t >> 6, c = t & 63;
            const long pr = prow[rr];
            float av = 0.f, bv = 0.f;
            if (pr >= 0) {
                const int mc = m0 + c;
                av = (mc < a.Ma ? a.A[pr * a.lda + mc] : (mc < Mt ? 1.f : 0.f)) * wrow[rr];
                bv = n0 + c < a.Nb ? a.B[pr * a.ldb + n0 + c] : 0.f;
            }
            As[rr][c] = av;
            Bs[rr][c] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { x[u] = As[k][ty * 4 + u]; y[u] = Bs[k][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(x[u], y[v], acc[u][v]);
        }
        __syncthreads();
    }
    float *P = a.part + (size_t)blockIdx.y * Mt * a.Nb;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int mm = m0 + ty * 4 + u, nn = n0 + tx * 4 + v;
            if (mm < Mt && nn < a.Nb) P[(size_t)mm * a.Nb + nn] = acc[u][v];
        }
}

// WIDE operands (both sides > 64 columns: the 256-wide layers, the h_el factor, the orbital layers): 128 x 128 tile per block, 8 x 8 outputs per
// thread (two 4-wide groups 64 apart on each side, so the float4 shared-memory reads of a quarter warp are contiguous), 16 rows per step, the next
// step's rows prefetched into registers while the current one is multiplied (one barrier per step).  FP32 CUDA cores: ~4 LDS.128 per 64 FMA.
__global__ void __launch_bounds__(256, 2) k_atb_wide(AtbArgs a) {
    __shared__ __align__(16) float As[2][16][128], Bs[2][16][128];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int Mt = a.Ma + a.ones;
    const int n_nb = (a.Nb + 127) / 128;
    const int m0 = (blockIdx.x / n_nb) * 128, n0 = (blockIdx.x % n_nb) * 128;
    const long r_begin = blockIdx.y * a.rows_per_split, r_end = min(a.rows, r_begin + a.rows_per_split);
    const bool vecA = (a.lda & 3) == 0 && ((size_t)a.A & 15) == 0, vecB = (a.ldb & 3) == 0 && ((size_t)a.B & 15) == 0;
    const int lrow = tid >> 5, lc = (tid & 31) * 4;              // loader: rows lrow and lrow + 8 of the step, columns lc .. lc + 3
    float acc[8][8] = {};
    float4 ra[2], rb[2];
    auto fetch = [&](long rc) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long r = rc + lrow + 8 * q;
            ra[q] = make_float4(0.f, 0.f, 0.f, 0.f); rb[q] = ra[q];
            if (r < r_end) {
                const unsigned r32 = (unsigned)r;
                const float w = a.wts ? a.wts[r32 / (unsigned)a.rpw] : 1.f;
                const long pr = a.rm.seg_len ? (long)(r32 / (unsigned)a.rm.seg_len) * a.rm.seg_stride + a.rm.seg_off + r32 % (unsigned)a.rm.seg_len : r;
                const int mc = m0 + lc, nc = n0 + lc;
                float va[4], vb[4];
                if (vecA && mc + 3 < a.Ma) { const float4 t = *reinterpret_cast<const float4 *>(a.A + pr * a.lda + mc); va[0] = t.x; va[1] = t.y; va[2] = t.z; va[3] = t.w; }
                else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) va[k] = mc + k < a.Ma ? a.A[pr * a.lda + mc + k] : (mc + k < Mt ? 1.f : 0.f);
                }
                if (vecB && nc + 3 < a.Nb) { const float4 t = *reinterpret_cast<const float4 *>(a.B + pr * a.ldb + nc); vb[0] = t.x; vb[1] = t.y; vb[2] = t.z; vb[3] = t.w; }
                else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) vb[k] = nc + k < a.Nb ? a.B[pr * a.ldb + nc + k] : 0.f;
                }
                ra[q] = make_float4(va[0] * w, va[1] * w, va[2] * w, va[3] * w);
                rb[q] = make_float4(vb[0], vb[1], vb[2], vb[3]);
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            *reinterpret_cast<float4 *>(&As[buf][lrow + 8 * q][lc]) = ra[q];
            *reinterpret_cast<float4 *>(&Bs[buf][lrow + 8 * q][lc]) = rb[q];
        }
    };
    fetch(r_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (long rc = r_begin; rc < r_end; rc += 16, buf ^= 1) {
        const bool more = rc + 16 < r_end;
        if (more) fetch(rc + 16);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]), a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]), b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
            const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, y[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(x[u], y[v], acc[u][v]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
    }
    float *P = a.part + (size_t)blockIdx.y * Mt * a.Nb;
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int mm = m0 + (u < 4 ? ty * 4 + u : 64 + ty * 4 + u - 4), nn = n0 + (v < 4 ? tx * 4 + v : 64 + tx * 4 + v - 4);
            if (mm < Mt && nn < a.Nb) P[(size_t)mm * a.Nb + nn] = acc[u][v];
        }
}

// The same product for NARROW operands (Ma, Nb <= 32: every pair / el-ion / ion layer): HBM-bound, ~(Ma + Nb) * 4 bytes and Mt * Nb FMAs per row.
// No shared-memory staging: each warp streams its own rows, every lane owns a 4 x 8 patch of the 32 x 32 product (+ its share of the ones row) and
// reads its 4 + 8 operands straight from the row (all lanes of a warp hit the same one or two 128-byte lines: broadcast loads).  The eight warps of
// a block are combined in shared memory in a fixed order.
__global__ void __launch_bounds__(256) k_atb_narrow(AtbArgs a) {
    __shared__ float red[8][33 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = (lane >> 2) * 4, n0 = (lane & 3) * 8;
    const int Mt = a.Ma + a.ones;
    const long r_begin = blockIdx.x * a.rows_per_split, r_end = min(a.rows, r_begin + a.rows_per_split);
    const bool vecA = a.Ma == 32 && (a.lda & 3) == 0 && ((size_t)a.A & 15) == 0;
    const bool vecB = a.Nb == 32 && (a.ldb & 3) == 0 && ((size_t)a.B & 15) == 0;
    const unsigned NN = (unsigned)(a.N * a.N);
    const float one_w = (lane >> 2) == 0 ? 1.f : 0.f;          // the ones row is accumulated by the lanes of the first patch row
    float acc[4][8] = {}, acc1[8] = {};
    constexpr int UNR = 4;
    for (long rb = r_begin + warp; rb < r_end; rb += 8 * UNR) {
        float av[UNR][4], bv[UNR][8], w[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long r = rb + 8 * u;
            bool ok = r < r_end;
            const unsigned r32 = (unsigned)r;                      // the launcher guarantees rows < 2^31: 32-bit divisions (the 64-bit ones cost more than the FMAs)
            if (ok && a.sel >= 0) {
                const unsigned pidx = r32 % NN, i = pidx / (unsigned)a.N, j = pidx - i * (unsigned)a.N;
                ok = ((((int)i < a.U) == ((int)j < a.U)) ? 0 : 1) == a.sel;
            }
            w[u] = ok ? (a.wts ? a.wts[r32 / (unsigned)a.rpw] : 1.f) : 0.f;
            const long pr = !ok ? -1 : (a.rm.seg_len ? (long)(r32 / (unsigned)a.rm.seg_len) * a.rm.seg_stride + a.rm.seg_off + r32 % (unsigned)a.rm.seg_len : r);
            if (pr >= 0 && vecA) {
                const float4 t = *reinterpret_cast<const float4 *>(a.A + pr * a.lda + m0);
                av[u][0] = t.x; av[u][1] = t.y; av[u][2] = t.z; av[u][3] = t.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) av[u][k] = (pr >= 0 && m0 + k < a.Ma) ? a.A[pr * a.lda + m0 + k] : 0.f;
            }
            if (pr >= 0 && vecB) {
                const float4 t0 = *reinterpret_cast<const float4 *>(a.B + pr * a.ldb + n0), t1 = *reinterpret_cast<const float4 *>(a.B + pr * a.ldb + n0 + 4);
                bv[u][0] = t0.x; bv[u][1] = t0.y; bv[u][2] = t0.z; bv[u][3] = t0.w; bv[u][4] = t1.x; bv[u][5] = t1.y; bv[u][6] = t1.z; bv[u][7] = t1.w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) bv[u][k] = (pr >= 0 && n0 + k < a.Nb) ? a.B[pr * a.ldb + n0 + k] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float x = av[u][k] * w[u];
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[k][v] = fmaf(x, bv[u][v], acc[k][v]);
            }
            const float x1 = w[u] * one_w;
#pragma unroll
            for (int v = 0; v < 8; ++v) acc1[v] = fmaf(x1, bv[u][v], acc1[v]);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int v = 0; v < 8; ++v)
            if (m0 + k < a.Ma && n0 + v < a.Nb) red[warp][(m0 + k) * a.Nb + n0 + v] = acc[k][v];
    if (a.ones && (lane >> 2) == 0)
#pragma unroll
        for (int v = 0; v < 8; ++v)
            if (n0 + v < a.Nb) red[warp][a.Ma * a.Nb + n0 + v] = acc1[v];
    __syncthreads();
    float *P = a.part + (size_t)blockIdx.x * Mt * a.Nb;
    for (int t = threadIdx.x; t < Mt * a.Nb; t += 256) {
        float sacc = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) sacc += red[wv][t];
        P[t] = sacc;
    }
}

// Second form of the narrow product: the per-row bookkeeping (row decode, walker weight, addresses) of k_atb_narrow costs more instructions than
// its 32 FMAs per lane, and the kernel is issue bound (ncu: 125-157 warp instructions per row, issue active 55-67 % at 25 % occupancy).  Here a
// HALF-warp takes a row and every lane an 8 x 8 patch: one pass of the loop body serves two rows, so the overhead per row halves while the loads in
// flight (2 rows x UNR) and the FMA count per row stay the same.  ONES / WTS are compile-time (the G factors need neither).
template <bool ONES, bool WTS>
__global__ void __launch_bounds__(256, 2) k_atb_narrow2(AtbArgs a) {
    __shared__ float red[8][33 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, hw = lane >> 4, l16 = lane & 15;
    const int m0 = (l16 >> 2) * 8, n0 = (l16 & 3) * 8;
    const int Mt = a.Ma + (ONES ? 1 : 0);
    const long r_begin = blockIdx.x * a.rows_per_split, r_end = min(a.rows, r_begin + a.rows_per_split);
    const bool vecA = a.Ma == 32 && (a.lda & 3) == 0 && ((size_t)a.A & 15) == 0;
    const bool vecB = a.Nb == 32 && (a.ldb & 3) == 0 && ((size_t)a.B & 15) == 0;
    const unsigned NN = (unsigned)(a.N * a.N);
    const bool first_m = (l16 >> 2) == 0;                  // these lanes also accumulate the ones row
    float acc[8][8] = {}, acc1[8] = {};
    constexpr int UNR = 2;
    for (long rb = r_begin + warp * 2 + hw; rb < r_end; rb += 16 * UNR) {
        float av[UNR][8], bv[UNR][8], w[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const long r = rb + 16 * u;
            bool ok = r < r_end;
            const unsigned r32 = (unsigned)r;
            if (ok && a.sel >= 0) {
                const unsigned pidx = r32 % NN, i = pidx / (unsigned)a.N, j = pidx - i * (unsigned)a.N;
                ok = ((((int)i < a.U) == ((int)j < a.U)) ? 0 : 1) == a.sel;
            }
            w[u] = ok ? (WTS ? a.wts[r32 / (unsigned)a.rpw] : 1.f) : 0.f;
            const long pr = !ok ? -1 : (a.rm.seg_len ? (long)(r32 / (unsigned)a.rm.seg_len) * a.rm.seg_stride + a.rm.seg_off + r32 % (unsigned)a.rm.seg_len : r);
            if (pr >= 0 && vecA) {
                const float4 t0 = *reinterpret_cast<const float4 *>(a.A + pr * a.lda + m0), t1 = *reinterpret_cast<const float4 *>(a.A + pr * a.lda + m0 + 4);
                av[u][0] = t0.x; av[u][1] = t0.y; av[u][2] = t0.z; av[u][3] = t0.w; av[u][4] = t1.x; av[u][5] = t1.y; av[u][6] = t1.z; av[u][7] = t1.w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) av[u][k] = (pr >= 0 && m0 + k < a.Ma) ? a.A[pr * a.lda + m0 + k] : 0.f;
            }
            if (pr >= 0 && vecB) {
                const float4 t0 = *reinterpret_cast<const float4 *>(a.B + pr * a.ldb + n0), t1 = *reinterpret_cast<const float4 *>(a.B + pr * a.ldb + n0 + 4);
                bv[u][0] = t0.x; bv[u][1] = t0.y; bv[u][2] = t0.z; bv[u][3] = t0.w; bv[u][4] = t1.x; bv[u][5] = t1.y; bv[u][6] = t1.z; bv[u][7] = t1.w;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) bv[u][k] = (pr >= 0 && n0 + k < a.Nb) ? a.B[pr * a.ldb + n0 + k] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float x = WTS ? av[u][k] * w[u] : av[u][k];          // (rows that do not count were loaded as zeros)
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[k][v] = fmaf(x, bv[u][v], acc[k][v]);
            }
            if (ONES) {
#pragma unroll
                for (int v = 0; v < 8; ++v) acc1[v] = fmaf(w[u], bv[u][v], acc1[v]);
            }
        }
    }
    // the two rows of a pass: half-warp 1 is added to half-warp 0 (fixed order)
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[k][v] += __shfl_down_sync(0xffffffffu, acc[k][v], 16);
    if (ONES) {
#pragma unroll
        for (int v = 0; v < 8; ++v) acc1[v] += __shfl_down_sync(0xffffffffu, acc1[v], 16);
    }
    if (hw == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int v = 0; v < 8; ++v)
                if (m0 + k < a.Ma && n0 + v < a.Nb) red[warp][(m0 + k) * a.Nb + n0 + v] = acc[k][v];
        if (ONES && first_m)
#pragma unroll
            for (int v = 0; v < 8; ++v)
                if (n0 + v < a.Nb) red[warp][a.Ma * a.Nb + n0 + v] = acc1[v];
    }
    __syncthreads();
    float *P = a.part + (size_t)blockIdx.x * Mt * a.Nb;
    for (int t = threadIdx.x; t < Mt * a.Nb; t += 256) {
        float sacc = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) sacc += red[wv][t];
        P[t] = sacc;
    }
}

// C[m, n] (ldc) = (accumulate ? C : 0) + scale * sum_split P[split][m, n]      (fixed summation order: deterministic)
// A block reduces 32 outputs: warp g sums the partials g, g + 8, ... of its 32 outputs, then the eight sub-sums are added in order.
__global__ void __launch_bounds__(256) k_atb_reduce(const float *__restrict__ part, int n_split, int Mt, int Nb, float *__restrict__ C, long ldc, float scale,
                                                     int accumulate) {
    __shared__ float sub[8][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long idx = blockIdx.x * 32L + lane, per = (long)Mt * Nb;
    float s = 0.f;
    if (idx < per)
        for (int k = g; k < n_split; k += 8) s += part[(size_t)k * per + idx];
    sub[g][lane] = s;
    __syncthreads();
    if (g == 0 && idx < per) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sub[k][lane];
        const int mm = (int)(idx / Nb), nn = (int)(idx - (long)mm * Nb);
        float *c = C + (long)mm * ldc + nn;
        *c = (accumulate ? *c : 0.f) + scale * t;
    }
}

// Operands of a WIDE product for the tensor-core path (launch_atb_tc, gemm_tc.cu): the reduction index r (walker x electron rows) becomes the
// contiguous one, through a 32 x 32 shared-memory transpose (coalesced on both sides).
//   mode 0:  At[s][m][k] = wrow(r) * Aaug[r, m],  r = s Kc + k   (FP32, chunk-major; the kernel splits it into tf32 halves in shared memory)
//   mode 1:  Bt_hi / Bt_lo[n][r] = tf32 halves of B[r, n]
// rows r >= rows (padding up to Rp = S Kc) are written as zeros.
__device__ __forceinline__ float to_tf32(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }
__global__ void __launch_bounds__(256) k_atb_prep(const float *__restrict__ src, long ld, int ncols, int ones, long rows, long Rp, int Kc, int rpw,
                                                  const float *__restrict__ wts, RowMap rm, int mode, float *__restrict__ o0, float *__restrict__ o1) {
    // (ncu on the first version: issue-bound, ~250 warp instructions per 32 x 32 tile, most of them 64-bit divisions of the row decode done by every
    // thread: the per-row work -- physical row, walker weight -- is done once per row by the first warp, and the indices are 32-bit)
    __shared__ float tile[32][33];
    __shared__ long poff[32];          // element offset of the row in src, or -1
    __shared__ float wrow[32];
    const unsigned r0 = blockIdx.x * 32u;                    // rows < 2^31 (launcher)
    const int c0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nc = ncols + ones;
    if (threadIdx.x < 32) {
        const unsigned r = r0 + threadIdx.x;
        const bool ok = r < (unsigned)rows;
        const long pr = !ok ? -1 : (rm.seg_len ? (long)(r / (unsigned)rm.seg_len) * rm.seg_stride + rm.seg_off + r % (unsigned)rm.seg_len : (long)r);
        poff[threadIdx.x] = ok ? pr * ld : -1;
        wrow[threadIdx.x] = ok ? (wts ? wts[r / (unsigned)rpw] : 1.f) : 0.f;
    }
    __syncthreads();
    const int c = c0 + tx;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int rr = ty + 8 * q;
        const long po = poff[rr];
        float v = 0.f;
        if (po >= 0 && c < nc) v = (c < ncols ? src[po + c] : 1.f) * wrow[rr];
        tile[rr][tx] = v;
    }
    __syncthreads();
    const unsigned r = r0 + tx, sc = r / (unsigned)Kc, k = r - sc * (unsigned)Kc;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int cc = ty + 8 * q, co = c0 + cc;
        if (co >= nc) continue;
        const float v = tile[tx][cc];
        if (mode == 0) o0[((size_t)sc * nc + co) * Kc + k] = v;
        else {
            const float h = to_tf32(v);
            const size_t o = (size_t)co * Rp + r;
            o0[o] = h;
            o1[o] = to_tf32(v - h);
        }
    }
}

struct GradCtx {
    dpe_model *m;
    cudaStream_t s;
    float *part; size_t part_floats;      // scratch of the split products
    const float *cot;                     // per-walker cotangents c_b of this chunk (gradient), or nullptr
    bool accumulate;                      // add to the outputs (second and later chunks)
    float *tcs; size_t tcs_floats;        // transposed operands of the wide products on the tensor-core path
};

// out[Ma + ones, Nb] (ldc) (+)= scale * Aaug^T diag(w) B
static int atb(GradCtx &g, float *out, long ldc, const float *A, long lda, int Ma, bool ones, const float *B, long ldb, int Nb, long rows, int rpw,
               const float *wts, float scale, int sel = -1, RowMap rm = RowMap{0, 0, 0}) {
    if (!out || rows <= 0) return DPE_OK;
    const dpe_dims &d = g.m->dims;
    AtbArgs a;
    a.A = A; a.lda = lda; a.Ma = Ma; a.ones = ones ? 1 : 0; a.B = B; a.ldb = ldb; a.Nb = Nb; a.rows = rows; a.rpw = rpw; a.wts = wts;
    a.sel = sel; a.N = d.n_el; a.U = d.n_up; a.rm = rm;
    const int Mt = Ma + a.ones;
    const size_t per = (size_t)Mt * Nb;
    if (Ma <= 32 && Nb <= 32 && rows < (1L << 31)) {               // narrow operands: the streaming kernel
        // two blocks are resident per SM (registers): one wave, and half the partial products of the earlier two-wave grid to reduce
        long n_blocks = std::min<long>((rows + 255) / 256, 148L * 2);
        while (n_blocks > 1 && per * n_blocks > g.part_floats) n_blocks = (n_blocks + 1) / 2;
        if (per * n_blocks > g.part_floats) return set_error(DPE_ERR_WORKSPACE, "gradient workspace too small for a %d x %d product", Mt, Nb);
        a.n_split = (int)n_blocks;
        a.rows_per_split = (rows + n_blocks - 1) / n_blocks;
        a.part = g.part;
        static const bool narrow_v1 = getenv("DPE_ATB_NARROW_V1") != nullptr;
        if (narrow_v1) k_atb_narrow<<<(unsigned)n_blocks, 256, 0, g.s>>>(a);
        else if (a.ones && a.wts) k_atb_narrow2<true, true><<<(unsigned)n_blocks, 256, 0, g.s>>>(a);
        else if (a.ones) k_atb_narrow2<true, false><<<(unsigned)n_blocks, 256, 0, g.s>>>(a);
        else if (a.wts) k_atb_narrow2<false, true><<<(unsigned)n_blocks, 256, 0, g.s>>>(a);
        else k_atb_narrow2<false, false><<<(unsigned)n_blocks, 256, 0, g.s>>>(a);
        DPE_LAUNCH_CHECK(g.m);
        k_atb_reduce<<<(int)((per + 31) / 32), 256, 0, g.s>>>(g.part, a.n_split, Mt, Nb, out, ldc, scale, g.accumulate ? 1 : 0);
        DPE_LAUNCH_CHECK(g.m);
        return DPE_OK;
    }
    static const bool atb_tc_off = getenv("DPE_ATB_TC") && atoi(getenv("DPE_ATB_TC")) == 0;
    if (Mt > 64 && Nb > 64 && rows >= 1024 && rows < (1L << 31) && sel < 0 && !(Nb & 3) && g.tcs && g.m->gemm_path == 1 && tc_pair_mode() >= 1 && !atb_tc_off) {
        // wide operands on the tensor cores: split-K over chunks of Kc rows, 3xTF32 (FP32-accurate) on the CTA-pair dense-layer kernel.
        // (Short products stay on the FP32 cores: the accumulation-bias compensation assumes full chunks, and there is nothing to gain.)
        // Kc balances waves of (256 feature x <= 256 row) tiles over the CTA pairs (74 on a B200) against the size of the partial products.
        const int tiles_min = (Mt + 255) / 256, nmma = (((Mt + tiles_min - 1) / tiles_min) + 31) / 32 * 32;
        const long tpc = (long)((Mt + nmma - 1) / nmma) * ((Nb + 255) / 256);
        int Kc = 0; double best = 0.0;
        for (int kc = 256; kc <= 2048; kc *= 2) {
            const long S = (rows + kc - 1) / kc;
            if (per * (size_t)S > g.part_floats || (size_t)(Mt + 2 * Nb) * (size_t)(S * kc) > g.tcs_floats) continue;
            const long n_pairs = std::max(1, g.m->n_sm / 2), waves = (S * tpc + n_pairs - 1) / n_pairs;
            const double cost = (double)waves * (kc * 120.0 + 25000.0) + (double)S * per * 8.0 / 6.0e12 * 1.9e9 / 1.0;     // clocks: MMA waves + partial write / read
            if (!Kc || cost < best) { Kc = kc; best = cost; }
        }
        if (Kc) {
            const long S = (rows + Kc - 1) / Kc, Rp = S * Kc;
            float *At = g.tcs, *Bh = At + (size_t)Mt * Rp, *Bl = Bh + (size_t)Nb * Rp;
            k_atb_prep<<<dim3((unsigned)(Rp / 32), (unsigned)((Mt + 31) / 32)), 256, 0, g.s>>>(A, lda, Ma, a.ones, rows, Rp, Kc, rpw, wts, rm, 0, At, nullptr);
            DPE_LAUNCH_CHECK(g.m);
            k_atb_prep<<<dim3((unsigned)(Rp / 32), (unsigned)((Nb + 31) / 32)), 256, 0, g.s>>>(B, ldb, Nb, 0, rows, Rp, Kc, 1, nullptr, rm, 1, Bh, Bl);
            DPE_LAUNCH_CHECK(g.m);
            const int e = launch_atb_tc(g.m, At, Bh, Bl, Rp, Kc, Mt, Nb, g.part, g.s);
            if (e == DPE_OK) {
                k_atb_reduce<<<(int)((per + 31) / 32), 256, 0, g.s>>>(g.part, (int)S, Mt, Nb, out, ldc, scale, g.accumulate ? 1 : 0);
                DPE_LAUNCH_CHECK(g.m);
                return DPE_OK;
            }
            if (e != DPE_ERR_UNSUPPORTED) return e;
        }
    }
    if (Mt > 64 && Nb > 64 && rows < (1L << 31) && sel < 0) {       // wide operands: 128 x 128 tiles
        const long tiles = (long)((Mt + 127) / 128) * ((Nb + 127) / 128);
        int n_split = (int)((rows + 255) / 256);
        while (n_split > 1 && tiles * n_split > 148L * 4) n_split = (n_split + 1) / 2;      // two blocks per SM, two waves
        while (n_split > 1 && per * n_split > g.part_floats) n_split = (n_split + 1) / 2;
        if (per * n_split > g.part_floats) return set_error(DPE_ERR_WORKSPACE, "gradient workspace too small for a %d x %d product", Mt, Nb);
        a.n_split = n_split;
        a.rows_per_split = ((rows + n_split - 1) / n_split + 15) / 16 * 16;
        a.part = g.part;
        k_atb_wide<<<dim3((unsigned)tiles, (unsigned)n_split), 256, 0, g.s>>>(a);
        DPE_LAUNCH_CHECK(g.m);
        k_atb_reduce<<<(int)((per + 31) / 32), 256, 0, g.s>>>(g.part, n_split, Mt, Nb, out, ldc, scale, g.accumulate ? 1 : 0);
        DPE_LAUNCH_CHECK(g.m);
        return DPE_OK;
    }
    const long tiles = (long)((Mt + 63) / 64) * ((Nb + 63) / 64);
    // every block walks its rows serially, 16 at a time: split the rows until ~8 blocks per SM are in flight, but keep >= 256 rows per block
    int n_split = (int)((rows + 255) / 256);
    while (n_split > 1 && tiles * n_split > 148L * 8) n_split = (n_split + 1) / 2;
    while (n_split > 1 && per * n_split > g.part_floats) n_split = (n_split + 1) / 2;
    if (per * n_split > g.part_floats) return set_error(DPE_ERR_WORKSPACE, "gradient workspace too small for a %d x %d product", Mt, Nb);
    a.n_split = n_split;
    a.rows_per_split = ((rows + n_split - 1) / n_split + 15) / 16 * 16;
    a.part = g.part;
    dim3 grid((unsigned)tiles, (unsigned)n_split);
    k_atb<<<grid, 256, 0, g.s>>>(a);
    DPE_LAUNCH_CHECK(g.m);
    k_atb_reduce<<<(int)((per + 31) / 32), 256, 0, g.s>>>(g.part, n_split, Mt, Nb, out, ldc, scale, g.accumulate ? 1 : 0);
    DPE_LAUNCH_CHECK(g.m);
    return DPE_OK;
}

// ------------------------------------------------------------------------------------------------ determinants: inverse
// FP64 Gauss-Jordan with partial pivoting, one block per (walker, determinant): Ainv [N, N] (FP32), log|det|, sign.
template <int T>
__global__ void __launch_bounds__(T) k_det_inverse(int N, int n_det, const float *__restrict__ mo, float *__restrict__ det, float *__restrict__ ainv) {
    // in place (no identity half), row interchanges undone as column interchanges in reverse order -- as k_det in orbitals_det.cu.
    // T = 32 (N <= 32): one warp per matrix, warp-level synchronisation only.
    extern __shared__ double aug[];          // [N][N + 1], then colp[N], then perm[N]
    __shared__ int piv_row;
    const int S = N + 1, tid = threadIdx.x;
    double *colp = aug + N * S;
    int *perm = reinterpret_cast<int *>(colp + N);
    auto sync = [] { if (T == 32) __syncwarp(); else __syncthreads(); };
    const long bd = blockIdx.x, b = bd / n_det;
    const int dt = (int)(bd - b * n_det), cols = n_det * N;
    const float *mob = mo + b * (long)N * cols + (long)dt * N;
    for (int e = tid; e < N * N; e += T) {
        const int i = e / N, o = e - i * N;
        aug[i * S + o] = (double)mob[(long)i * cols + o];
    }
    sync();
    LogDetAcc logdet;
    float sign = 1.f;
    for (int p = 0; p < N; ++p) {
        if (tid < 32) {
            double best = -1.0; int bi = p;
            for (int i = p + tid; i < N; i += 32) { const double v = fabs(aug[i * S + p]); if (v > best) { best = v; bi = i; } }
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) { piv_row = bi; perm[p] = bi; }
        }
        sync();
        const int pr = piv_row;
        if (pr != p) {
            for (int o = tid; o < N; o += T) { const double t = aug[p * S + o]; aug[p * S + o] = aug[pr * S + o]; aug[pr * S + o] = t; }
            sign = -sign;
            sync();
        }
        const double piv = aug[p * S + p];
        logdet.mul(piv);
        if (piv < 0.0) sign = -sign;
        const double inv = 1.0 / piv;
        for (int i = tid; i < N; i += T) colp[i] = aug[i * S + p];
        sync();
        for (int o = tid; o < N; o += T) aug[p * S + o] = o == p ? inv : aug[p * S + o] * inv;
        sync();
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
        sync();
    }
    if (tid == 0) {
        det[bd * 2] = (float)logdet.value(); det[bd * 2 + 1] = sign;
        int pv[64];
        for (int k = 0; k < N; ++k) pv[k] = perm[k];
        for (int k = 0; k < N; ++k) perm[k] = k;
        for (int pp = N - 1; pp >= 0; --pp) { const int t = perm[pp]; perm[pp] = perm[pv[pp]]; perm[pv[pp]] = t; }      // perm becomes the column index map
    }
    sync();
    for (int e = tid; e < N * N; e += T) {
        const int q = e / N, i = e - q * N;
        ainv[bd * (long)N * N + e] = (float)aug[q * S + perm[i]];          // Ainv[q][i]
    }
}

// The same inverse with ONE ROW PER THREAD held in registers (N <= NP <= 64; T = 32 or 64 threads per matrix, four matrices per block for T = 32).
// The shared-memory form above moves ~1.2 MB through shared memory per 42 x 42 matrix (every element read and written once per pivot); here a
// pivot step costs one warp-wide max (REDUX on a magnitude | row key), one row written to shared memory by its owner and N broadcast LDS.64 by the
// others.  Rows are never exchanged: column p takes its pivot from the not-yet-used thread t_p with the largest |a[p]|, which is the row-swapped
// algorithm with logical row p living in thread t_p; at the end  Ainv[p][t_k] = a_{t_p}[k]  and the sign carries the parity of k -> t_k.
template <int NP, int T>
__global__ void __launch_bounds__(T == 32 ? 128 : 64) k_det_inverse_rows(int N, int n_det, long n_mat, const float *__restrict__ mo, float *__restrict__ det,
                                                                          float *__restrict__ ainv) {
    constexpr int MPB = T == 32 ? 4 : 1;
    __shared__ double rowp[MPB][NP], pv[MPB][NP];
    __shared__ unsigned wkey[MPB][2];
    __shared__ int tk[MPB][NP];
    const int sub = threadIdx.x / T, t = threadIdx.x % T;
    const long bd = blockIdx.x * (long)MPB + sub;
    if (bd >= n_mat) return;                              // whole warps (T = 32) or the whole block (T = 64)
    auto sync = [] { if (T == 32) __syncwarp(); else __syncthreads(); };
    const long b = bd / n_det;
    const int dt = (int)(bd - b * n_det), cols = n_det * N;
    const float *mob = mo + b * (long)N * cols + (long)dt * N;
    double a[NP];
#pragma unroll
    for (int o = 0; o < NP; ++o) a[o] = (t < N && o < N) ? (double)mob[(long)t * cols + o] : 0.0;
    bool used = t >= N;
    int my_p = -1;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (p < N) {
            // key: [31] row still free, [30:6] exponent and 14 leading mantissa bits of |a[p]| (FP64), [5:0] 63 - row (ties go to the lowest row)
            const unsigned key = used ? 0u : (0x80000000u | ((unsigned)__double2hiint(fabs(a[p])) & 0x7fffffc0u) | (unsigned)(63 - t));
            unsigned best = __reduce_max_sync(0xffffffffu, key);
            if (T == 64) {
                if ((threadIdx.x & 31) == 0) wkey[sub][threadIdx.x >> 5] = best;
                __syncthreads();
                best = max(wkey[sub][0], wkey[sub][1]);
            }
            const int tp = 63 - (int)(best & 63u);
            if (t == tp) {
                const double piv = a[p], inv = 1.0 / piv;
#pragma unroll
                for (int o = 0; o < NP; ++o)
                    if (o < N) { a[o] = o == p ? inv : a[o] * inv; rowp[sub][o] = a[o]; }
                pv[sub][p] = piv;
                tk[sub][p] = t;
                used = true;
                my_p = p;
            }
            sync();
            if (t != tp) {
                const double f = a[p];
#pragma unroll
                for (int o = 0; o < NP; ++o)
                    if (o < N) a[o] = o == p ? -f * rowp[sub][p] : fma(-f, rowp[sub][o], a[o]);
            }
            sync();                                       // the next pivot row overwrites rowp (and wkey)
        }
    }
    if (my_p >= 0) {
        float *out = ainv + bd * (long)N * N + (long)my_p * N;
#pragma unroll
        for (int k = 0; k < NP; ++k)
            if (k < N) out[tk[sub][k]] = (float)a[k];
    }
    if (t == 0) {
        LogDetAcc logdet;
        float sign = 1.f;
        unsigned long long seen = 0ull;
        for (int k = 0; k < N; ++k) {
            const double piv = pv[sub][k];
            logdet.mul(piv);
            if (piv < 0.0) sign = -sign;
            if (!((seen >> k) & 1ull)) {                  // cycle of the permutation k -> t_k: a cycle of length L contributes (-1)^(L - 1)
                int j = k, len = 0;
                while (!((seen >> j) & 1ull)) { seen |= 1ull << j; j = tk[sub][j]; ++len; }
                if (!(len & 1)) sign = -sign;
            }
        }
        det[bd * 2] = (float)logdet.value(); det[bd * 2 + 1] = sign;
    }
}

// coef[b, d] = d log psi^2 / d log|det_d| = 2 rho w_d   (wavefunction.py:77-83; the shift is the constant max)
__global__ void k_bw_combine(int Bc, int n_det, const float *__restrict__ det, float *__restrict__ coef, float *__restrict__ logpsi2) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= Bc) return;
    const float *db = det + (long)b * n_det * 2;
    float best = -INFINITY;
    for (int d = lane; d < n_det; d += 32) best = fmaxf(best, db[d * 2]);
    for (int o = 16; o; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    double psi = 0.0;
    for (int d = lane; d < n_det; d += 32) psi += (double)(db[d * 2 + 1] * expf(db[d * 2] - best));
    for (int o = 16; o; o >>= 1) psi += __shfl_xor_sync(0xffffffffu, psi, o);
    const double apsi = fabs(psi), rho = apsi / (apsi + 1e-8);
    for (int d = lane; d < n_det; d += 32) coef[(long)b * n_det + d] = (float)(2.0 * rho * (double)(db[d * 2 + 1] * expf(db[d * 2] - best)) / psi);
    if (lane == 0 && logpsi2) logpsi2[b] = (float)(2.0 * (log(apsi + 1e-8) + (double)best));
}

// ------------------------------------------------------------------------------------------------ orbitals backward
// thread = orbital column, loops over the (walker, electron) rows of its split:
//   dmo = coef[b, det] Ainv[b, det][orb][i];  env and its parameter derivatives recomputed;  dbf = dmo env;  denv = dmo bf
//   d weights[J, col] += c_b denv exp(-a d);  d alpha[J, col] += c_b denv w exp(-a d) (-d) sigmoid(alpha)
// partial sums per split: part[split][spin][2][I][cols]
template <int MAXI>          // ions per register block (4 / 8 / 16): four arrays of MAXI floats per thread decide the occupancy of this latency-bound kernel
__global__ void __launch_bounds__(64) k_bw_orbitals(int Bc, int N, int U, int I, int n_det, const float *__restrict__ r, const float *__restrict__ R,
                                                    const float *__restrict__ coef, const float *__restrict__ ainv, const float *__restrict__ bf,
                                                    const float *__restrict__ spa_up, const float *__restrict__ spa_dn, const float *__restrict__ alpha_up,
                                                    const float *__restrict__ alpha_dn, const float *__restrict__ w_up, const float *__restrict__ w_dn,
                                                    const float *__restrict__ cot, float *__restrict__ dbf, float *__restrict__ part, int walkers_per_split) {
    // thread = orbital column, block row = a few walkers (the launcher sizes the split so that ~8 blocks per SM exist even for small batches)
    const int cols = n_det * N;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= cols) return;
    const int dt = col / N, q = col - dt * N;
    const int b0 = blockIdx.y * walkers_per_split, b1 = min(Bc, b0 + walkers_per_split);
    for (int sp = 0; sp < 2; ++sp) {
        const float *spa = sp ? spa_dn : spa_up, *al = sp ? alpha_dn : alpha_up, *wt = sp ? w_dn : w_up;
        const int i_lo = sp ? U : 0, i_hi = sp ? N : U;
        for (int J0 = 0; J0 < I; J0 += MAXI) {
            const int nJ = min(MAXI, I - J0);
            const bool one_block = I <= MAXI;                 // all ions in this block: the envelope itself comes out of the same loop
            float gw[MAXI], ga[MAXI], sa[MAXI], sw[MAXI];
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                gw[k] = 0.f; ga[k] = 0.f;
                sa[k] = k < nJ ? spa[(long)(J0 + k) * cols + col] : 0.f;
                sw[k] = k < nJ ? wt[(long)(J0 + k) * cols + col] : 0.f;
            }
            for (int b = b0; b < b1; ++b) {
                const float cb = cot ? cot[b] : 1.f;
                const float cf = coef[(long)b * n_det + dt];
                for (int i = i_lo; i < i_hi; ++i) {
                    const long row = (long)b * N + i;
                    const float dmo = cf * ainv[((long)b * n_det + dt) * N * N + (long)q * N + i];
                    const float *ri = r + row * 3;
                    const float rx = ri[0], ry = ri[1], rz = ri[2];
                    float env = 0.f;
                    const float denv = dmo * bf[row * cols + col];
                    if (J0 == 0 && !one_block) {
                        for (int J = 0; J < I; ++J) {
                            const float dx = rx - R[J * 3], dy = ry - R[J * 3 + 1], dz = rz - R[J * 3 + 2];
                            env += wt[(long)J * cols + col] * expf(-spa[(long)J * cols + col] * sqrtf(dx * dx + dy * dy + dz * dz));
                        }
                        dbf[row * cols + col] = dmo * env;
                    }
#pragma unroll
                    for (int k = 0; k < MAXI; ++k)
                        if (k < nJ) {
                            const int J = J0 + k;
                            const float dx = rx - R[J * 3], dy = ry - R[J * 3 + 1], dz = rz - R[J * 3 + 2];
                            const float dd = sqrtf(dx * dx + dy * dy + dz * dz);
                            const float e = expf(-sa[k] * dd);
                            env = fmaf(sw[k], e, env);
                            gw[k] = fmaf(cb * denv, e, gw[k]);
                            ga[k] = fmaf(cb * denv, -dd * sw[k] * e, ga[k]);
                        }
                    if (one_block) dbf[row * cols + col] = dmo * env;
                }
            }
            for (int k = 0; k < nJ; ++k) {
                const int J = J0 + k;
                const float a0 = al[(long)J * cols + col];
                const float sig = 1.f / (1.f + expf(-a0));                     // d softplus / d alpha
                float *P = part + ((((size_t)blockIdx.y * 2 + sp) * 2) * I) * cols;
                P[(size_t)J * cols + col] = gw[k];
                P[((size_t)I + J) * cols + col] = ga[k] * sig;
            }
        }
    }
}

// Transferable-atomic-orbital head (orbitals_det.cu: k_tao_orbitals), backward of the value channel:
//   mo[b, i, col] = sum_J g[b, i, J, col] exp(-x_s[J, col] |r_i - R_J|)   =>   dg[b, i, J, col] = dmo[b, i, col] exp(-x_s[J, col] |r_i - R_J|),
//   dmo[b, i, (d, q)] = coef[b, d] Ainv_d[q, i].   g is overwritten by dg.  (The backflow matrix and the exponents come from the geometry cache:
//   their cotangents belong to the geometry-only nets, which are outside this library.)
__global__ void __launch_bounds__(256) k_bw_tao(long total, int N, int U, int I, int n_det, const float *__restrict__ r, const float *__restrict__ R,
                                                const float *__restrict__ coef, const float *__restrict__ ainv, const float *__restrict__ ex_same,
                                                const float *__restrict__ ex_diff, float *__restrict__ g) {
    const int cols = n_det * N;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int col = (int)(idx % cols);
        const long bi = idx / cols;
        const long b = bi / N;
        const int i = (int)(bi - b * N), dt = col / N, q = col - dt * N;
        const float dmo = coef[b * n_det + dt] * ainv[(b * n_det + dt) * N * N + (long)q * N + i];
        const float *ex = ((i < U) == (q < U)) ? ex_same : ex_diff;
        const float *ri = r + bi * 3;
        float *gp = g + bi * (long)I * cols + col;
        for (int J = 0; J < I; ++J) {
            const float dx = ri[0] - R[J * 3], dy = ri[1] - R[J * 3 + 1], dz = ri[2] - R[J * 3 + 2];
            gp[(long)J * cols] = dmo * expf(-ex[(long)J * cols + col] * sqrtf(dx * dx + dy * dy + dz * dz));
        }
    }
}

// leaf[J, col] (+)= sum_split part[split][spin][which][J][col]
__global__ void __launch_bounds__(256) k_bw_env_reduce(const float *__restrict__ part, int n_split, int I, int cols, int sp, int which, float *__restrict__ out,
                                                       int accumulate) {
    // a block reduces 32 outputs: warp g sums the splits g, g + 8, ..., then the eight sub-sums are added in order (deterministic)
    __shared__ float sub[8][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long idx = blockIdx.x * 32L + lane, n = (long)I * cols;
    float s = 0.f;
    if (idx < n)
        for (int k = g; k < n_split; k += 8) s += part[((((size_t)k * 2 + sp) * 2 + which) * I) * cols + idx];
    sub[g][lane] = s;
    __syncthreads();
    if (g == 0 && idx < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sub[k][lane];
        out[idx] = (accumulate ? out[idx] : 0.f) + t;
    }
}

// ------------------------------------------------------------------------------------------------ dense layers backward (elementwise parts)
// dz = dy (1 - y^2) for rows [Bc N]; sumdz[b] = sum_i dz[b, i]
__global__ void __launch_bounds__(256) k_bw_tanh(int N, int width, const float *__restrict__ dy, long lddy, const float *__restrict__ y, long ldy,
                                                 float *__restrict__ dz, float *__restrict__ sumdz) {
    const long b = blockIdx.x;
    for (int f = threadIdx.x; f < width; f += blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < N; ++i) {
            const long row = b * N + i;
            const float yy = y[row * ldy + f];
            const float v = dy[row * lddy + f] * (1.f - yy * yy);
            dz[row * width + f] = v;
            s += v;
        }
        if (sumdz) sumdz[b * width + f] = s;
    }
}

// sumx[b, :] = sum_i x[b, i, :width]  (+ a trailing column N: the ones column of the layer input summed over electrons)
__global__ void __launch_bounds__(256) k_sum_electrons(int N, int width, const float *__restrict__ x, long ldx, float *__restrict__ sumx) {
    const long b = blockIdx.x;
    for (int f = threadIdx.x; f <= width; f += blockDim.x) {
        float s = 0.f;
        if (f < width) for (int i = 0; i < N; ++i) s += x[(b * N + i) * ldx + f];
        else s = (float)N;
        sumx[b * (width + 1) + f] = s;
    }
}

// SchNet convolution backward (ferminet_embedding.py:159-168):  conv_ee[i] = sum_j w_ij hm_j
//   dzw[b, i, j, f] = dcee[b, i, f] hm[b, j, f] (1 - w_ij^2)      dhm[b, j, f] = sum_i w[b, i, j, f] dcee[b, i, f];  dzhm = dhm (1 - hm^2)
__global__ void __launch_bounds__(256) k_bw_conv(int N, int emb, const float *__restrict__ dx, long lddx, int col_ee, const float *__restrict__ hm,
                                                 const float *__restrict__ pw, float *__restrict__ dzw, float *__restrict__ dzhm) {
    const long b = blockIdx.x;
    for (int t = threadIdx.x; t < N * emb; t += blockDim.x) {
        const int j = t / emb, f = t - j * emb;
        const float h = hm[(b * N + j) * emb + f];
        float s = 0.f;
        for (int i = 0; i < N; ++i) {
            const float dc = dx[(b * N + i) * lddx + col_ee + f];
            const long pidx = ((b * N + i) * N + j) * (long)emb + f;
            const float w = pw[pidx];
            dzw[pidx] = dc * h * (1.f - w * w);
            s = fmaf(w, dc, s);
        }
        dzhm[(b * N + j) * emb + f] = s * (1.f - h * h);
    }
}

// dy_prev[b, i, :d_in] = dx[b, i, :d_in] + dhmap[b, i, :] + dmean[b, spin(i) block] / n_spin;   dcei[b, i, :] = dx[b, i, d_in + emb : k_main]
__global__ void __launch_bounds__(256) k_bw_gather(int N, int U, int d_in, int emb, int dE, const float *__restrict__ dx, long lddx,
                                                   const float *__restrict__ dhmap, const float *__restrict__ dmean, float *__restrict__ dy_prev,
                                                   long lddy, float *__restrict__ dcei) {
    const long row = blockIdx.x;
    const long b = row / N;
    const int i = (int)(row - b * N);
    const float inv = i < U ? 1.f / (float)U : 1.f / (float)(N - U);
    const float *dm = dmean + b * 2 * d_in + (i < U ? 0 : d_in);
    if (dy_prev)
        for (int f = threadIdx.x; f < d_in; f += blockDim.x) dy_prev[row * lddy + f] = dx[row * lddx + f] + dhmap[row * d_in + f] + dm[f] * inv;
    for (int f = threadIdx.x; f < dE; f += blockDim.x) dcei[row * dE + f] = dx[row * lddx + d_in + emb + f];
}

// ------------------------------------------------------------------------------------------------ pair stream backward
// One warp per ORDERED pair (i, j); lane = feature.  Recomputes the chain x^0 = d_ij, x^{it+1} = res(tanh(x^it W_h + b), x^it) and
// w^it = tanh(x^it W_w + b), then walks back with the cotangents dzw^it of the w layers (k_bw_conv):
//   dx^it = dzw^it W_w^T + dzh^it W_h^T + (residual ? dx^{it+1} / sqrt 2 : 0),   dzh^it = dx^{it+1} (res ? 1/sqrt 2 : 1) (1 - t^2)
// Outputs for the products: px[it] = x^it (layer input), dzh[it] and a copy of dzw[it], all three with the rows ordered class-major
// ([walker][same-spin pair] then [walker][different-spin pair]): every product of a w_same / w_diff / h_same / h_diff layer then streams a
// contiguous row range instead of filtering every second row away.
struct PairBwArgs {
    const float *r;
    const float *ww[DPE_MAX_ITER][2], *wb[DPE_MAX_ITER][2], *hw[DPE_MAX_ITER][2], *hb[DPE_MAX_ITER][2];
    const float *dzw[DPE_MAX_ITER];
    float *px[DPE_MAX_ITER], *dzh[DPE_MAX_ITER], *dzw_cm[DPE_MAX_ITER];     // written CLASS-MAJOR: [same-spin pairs of all walkers | different-spin pairs]
    int dP[DPE_MAX_ITER];
    int n_iter, N, U, emb;
    long n_pairs;
};

// The small dense layers of the pair / el-ion streams are staged in shared memory with row stride 33: W[k, n] at k * 33 + n, so both the forward
// product (lane = n, bank n + k) and the transposed one (lane = k, bank k + n) are bank-conflict free.
constexpr int WS = 33, WMAT = 32 * WS;
__device__ __forceinline__ void stage_weight(float *dst, const float *__restrict__ W, int din, int dout) {
    for (int t = threadIdx.x; t < din * dout; t += blockDim.x) dst[(t / dout) * WS + (t % dout)] = W[t];
}
__device__ __forceinline__ float warp_matvec(const float *W, int din, int dout, float x, int lane) {
    // y[lane] = sum_k x[k] W[k, lane]     (x distributed over lanes)
    float y = 0.f;
    for (int k = 0; k < din; ++k) y = fmaf(__shfl_sync(0xffffffffu, x, k), W[k * WS + lane], y);
    return lane < dout ? y : 0.f;
}
__device__ __forceinline__ float warp_matvec_t(const float *W, int din, int dout, float dy, int lane) {
    // dx[lane] = sum_n dy[n] W[lane, n]    (dy distributed over lanes)
    float dx = 0.f;
    for (int n = 0; n < dout; ++n) dx = fmaf(__shfl_sync(0xffffffffu, dy, n), W[lane * WS + n], dx);
    return lane < din ? dx : 0.f;
}

__global__ void __launch_bounds__(256) k_bw_pair(PairBwArgs a) {
    extern __shared__ float wsm[];                    // [it][class][w | h][32 x 33], zero outside the layer shapes
    for (int t = threadIdx.x; t < a.n_iter * 4 * WMAT; t += blockDim.x) wsm[t] = 0.f;
    __syncthreads();
    for (int it = 0; it < a.n_iter; ++it)
        for (int c = 0; c < 2; ++c) {
            stage_weight(wsm + ((it * 2 + c) * 2) * WMAT, a.ww[it][c], a.dP[it], a.emb);
            if (it + 1 < a.n_iter) stage_weight(wsm + ((it * 2 + c) * 2 + 1) * WMAT, a.hw[it][c], a.dP[it], a.dP[it + 1]);
        }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (long p = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; p < a.n_pairs; p += warps_total) {
    const int N = a.N;
    const long b = p / ((long)N * N);
    const int ij = (int)(p - b * N * N), i = ij / N, j = ij - i * N;
    const int sd = ((i < a.U) == (j < a.U)) ? 0 : 1;
    const int U = a.U, D = N - U, n_same = U * U + D * D, n_diff = 2 * U * D;
    const long n_walkers = a.n_pairs / ((long)N * N);
    const long pc = sd == 0 ? b * n_same + (i < U ? i * U + j : U * U + (i - U) * D + (j - U))
                            : n_walkers * n_same + b * n_diff + (i < U ? i * D + (j - U) : U * D + (i - U) * U + j);       // class-major row
    const float *ri = a.r + (b * N + i) * 3, *rj = a.r + (b * N + j) * 3;
    const float dx0 = rj[0] - ri[0], dy0 = rj[1] - ri[1], dz0 = rj[2] - ri[2];
    const float dist = i == j ? 0.f : sqrtf(dx0 * dx0 + dy0 * dy0 + dz0 * dz0);
    float x[DPE_MAX_ITER], t[DPE_MAX_ITER];          // layer inputs and tanh outputs of the h layers, this lane's feature
    x[0] = lane == 0 ? dist : 0.f;
#pragma unroll
    for (int it = 0; it < DPE_MAX_ITER; ++it) {
        if (it < a.n_iter) {
            if (lane < a.dP[it]) a.px[it][pc * a.dP[it] + lane] = x[it];
            if (it + 1 < a.n_iter) {
                const int dn = a.dP[it + 1];
                const float z = warp_matvec(wsm + ((it * 2 + sd) * 2 + 1) * WMAT, a.dP[it], dn, x[it], lane) + (lane < dn ? a.hb[it][sd][lane] : 0.f);
                t[it] = lane < dn ? tanhf(z) : 0.f;
                x[it + 1] = a.dP[it] == dn ? (x[it] + t[it]) * 0.70710678118654752f : t[it];
            }
        }
    }
    float dxn = 0.f;                                  // cotangent of x^{it+1}
#pragma unroll
    for (int it = DPE_MAX_ITER - 1; it >= 0; --it) {
        if (it < a.n_iter) {
            const float dzw = lane < a.emb ? a.dzw[it][p * a.emb + lane] : 0.f;
            if (lane < a.emb) a.dzw_cm[it][pc * a.emb + lane] = dzw;
            float dxc = warp_matvec_t(wsm + ((it * 2 + sd) * 2) * WMAT, a.dP[it], a.emb, dzw, lane);
            if (it + 1 < a.n_iter) {
                const int dn = a.dP[it + 1];
                const bool res = a.dP[it] == dn;
                const float dzh = dxn * (res ? 0.70710678118654752f : 1.f) * (1.f - t[it] * t[it]);
                if (lane < dn) a.dzh[it][pc * dn + lane] = dzh;
                dxc += warp_matvec_t(wsm + ((it * 2 + sd) * 2 + 1) * WMAT, a.dP[it], dn, dzh, lane);
                if (res) dxc += dxn * 0.70710678118654752f;
            }
            dxn = dxc;
        }
    }
    }
}

// The same backward for the default widths (pair features 1 -> 32 -> 32 -> ..., 32 convolution features): ONE THREAD per ordered pair.  The warp-per-pair
// kernel above spends its time in the MIO pipe (32 SHFL + 32 LDS per 32 x 32 product and pair); with a row per thread the operand vector lives in
// registers and every lane of a warp reads the SAME weight row, so a product costs 8 broadcast LDS.128 per pair and the kernel is FMA / HBM bound.
// The layer inputs x^it go to global memory on the way forward (they are outputs anyway) and are read back on the way down; the tanh output of a
// residual layer is recovered as t = sqrt2 x^{it+1} - x^it.
constexpr int PBR_THREADS = 384;          // one block per SM: one copy of the weights (64 KB), the rest of the 256 KB stays L1 for the row-per-thread global accesses
__device__ __forceinline__ void st_row32(float *dst, const float (&v)[32]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) reinterpret_cast<float4 *>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
// acc[k] += sum_n v[n] W[k, n] for the four n of one float4 of v   (W row-major [32][32] in shared memory, all lanes read the same address)
__device__ __forceinline__ void acc_wt4(float (&acc)[32], const float *W, int n4, const float4 v) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const float4 w = *reinterpret_cast<const float4 *>(W + k * 32 + n4 * 4);
        acc[k] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[k]))));
    }
}
__global__ void __launch_bounds__(PBR_THREADS) k_bw_pair_rows(PairBwArgs a) {
    extern __shared__ __align__(16) float wsm[];      // [it][class][w | h][32 x 32], then the h biases [it][class][32]
    const int nit = a.n_iter;
    float *bsm = wsm + nit * 4 * 1024;
    for (int t = threadIdx.x; t < nit * 4 * 1024; t += blockDim.x) {
        const int e = t & 1023, wh = (t >> 10) & 1, c = (t >> 11) & 1, it = t >> 12, k = e >> 5, n = e & 31;
        float v = 0.f;
        if (wh == 0) { if (k < a.dP[it]) v = a.ww[it][c][k * 32 + n]; }
        else if (it + 1 < nit && k < a.dP[it]) v = a.hw[it][c][k * 32 + n];
        wsm[t] = v;
    }
    for (int t = threadIdx.x; t < nit * 64; t += blockDim.x) {
        const int n = t & 31, c = (t >> 5) & 1, it = t >> 6;
        bsm[t] = it + 1 < nit ? a.hb[it][c][n] : 0.f;
    }
    __syncthreads();
    const int N = a.N, U = a.U, D = N - U, n_same = U * U + D * D, n_diff = 2 * U * D;
    const long n_walkers = a.n_pairs / ((long)N * N);
    constexpr float RS2 = 0.70710678118654752f, S2 = 1.41421356237309505f;
    // threads walk the pairs in CLASS-MAJOR order (the order of px / dzh / dzw_cm): all lanes of a warp but one boundary warp use the same
    // weight matrices, so the weight rows are true broadcasts (in pair order every warp mixes the two spin classes: two wavefronts per LDS)
    const long n_same_rows = n_walkers * n_same;
    for (long pc = blockIdx.x * (long)blockDim.x + threadIdx.x; pc < a.n_pairs; pc += (long)gridDim.x * blockDim.x) {
        const int sd = pc < n_same_rows ? 0 : 1;
        long b; int i, j;
        if (sd == 0) {
            b = pc / n_same;
            const int k = (int)(pc - b * n_same);
            if (k < U * U) { i = k / U; j = k - i * U; }
            else { const int k2 = k - U * U; i = U + k2 / D; j = U + (k2 - (i - U) * D); }
        } else {
            const long q = pc - n_same_rows;
            b = q / n_diff;
            const int k = (int)(q - b * n_diff);
            if (k < U * D) { i = k / D; j = U + (k - i * D); }
            else { const int k2 = k - U * D; i = U + k2 / U; j = k2 - (i - U) * U; }
        }
        const long p = (b * N + i) * N + j;                  // row of dzw (pair order)
        const float *ri = a.r + (b * N + i) * 3, *rj = a.r + (b * N + j) * 3;
        const float dx0 = rj[0] - ri[0], dy0 = rj[1] - ri[1], dz0 = rj[2] - ri[2];
        const float dist = i == j ? 0.f : sqrtf(dx0 * dx0 + dy0 * dy0 + dz0 * dz0);
        // ---- forward: x^1 = tanh(d W_h^0 + b), x^{it+1} = (x^it + tanh(x^it W_h^it + b)) / sqrt2; every x^it is stored (class-major)
        a.px[0][pc] = dist;
        float x[32];
        {
            const float *W = wsm + (sd * 2 + 1) * 1024, *bb = bsm + sd * 32;
#pragma unroll
            for (int n = 0; n < 32; ++n) x[n] = tanhf(fmaf(dist, W[n], bb[n]));
        }
        st_row32(a.px[1] + pc * 32, x);
        for (int it = 1; it + 1 < nit; ++it) {
            const float *W = wsm + ((it * 2 + sd) * 2 + 1) * 1024, *bb = bsm + (it * 2 + sd) * 32;
            float z[32];
#pragma unroll
            for (int n = 0; n < 32; ++n) z[n] = bb[n];
#pragma unroll
            for (int k = 0; k < 32; ++k)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = *reinterpret_cast<const float4 *>(W + k * 32 + q * 4);
                    z[4 * q] = fmaf(x[k], w.x, z[4 * q]); z[4 * q + 1] = fmaf(x[k], w.y, z[4 * q + 1]);
                    z[4 * q + 2] = fmaf(x[k], w.z, z[4 * q + 2]); z[4 * q + 3] = fmaf(x[k], w.w, z[4 * q + 3]);
                }
#pragma unroll
            for (int n = 0; n < 32; ++n) x[n] = (x[n] + tanhf(z[n])) * RS2;
            st_row32(a.px[it + 1] + pc * 32, x);
        }
        // ---- backward: dx^it = dzw^it W_w^T + dzh^it W_h^T + dx^{it+1} / sqrt2,  dzh^it = dx^{it+1} / sqrt2 (1 - t^2)
        float dxn[32];
#pragma unroll
        for (int n = 0; n < 32; ++n) dxn[n] = 0.f;
        for (int it = nit - 1; it >= 1; --it) {
            float dxc[32];
            const bool has_h = it + 1 < nit;
#pragma unroll
            for (int n = 0; n < 32; ++n) dxc[n] = has_h ? dxn[n] * RS2 : 0.f;
            const float *Ww = wsm + ((it * 2 + sd) * 2) * 1024, *Wh = Ww + 1024;
            const float4 *gz = reinterpret_cast<const float4 *>(a.dzw[it] + p * 32);
            float4 *gcm = reinterpret_cast<float4 *>(a.dzw_cm[it] + pc * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = gz[q];
                gcm[q] = v;
                acc_wt4(dxc, Ww, q, v);
            }
            if (has_h) {
                const float4 *x0 = reinterpret_cast<const float4 *>(a.px[it] + pc * 32), *x1 = reinterpret_cast<const float4 *>(a.px[it + 1] + pc * 32);
                float4 *gh = reinterpret_cast<float4 *>(a.dzh[it] + pc * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 u0 = x0[q], u1 = x1[q];
                    const float t0 = fmaf(S2, u1.x, -u0.x), t1 = fmaf(S2, u1.y, -u0.y), t2 = fmaf(S2, u1.z, -u0.z), t3 = fmaf(S2, u1.w, -u0.w);
                    const float4 v = make_float4(dxn[4 * q] * RS2 * (1.f - t0 * t0), dxn[4 * q + 1] * RS2 * (1.f - t1 * t1),
                                                 dxn[4 * q + 2] * RS2 * (1.f - t2 * t2), dxn[4 * q + 3] * RS2 * (1.f - t3 * t3));
                    gh[q] = v;
                    acc_wt4(dxc, Wh, q, v);
                }
            }
#pragma unroll
            for (int n = 0; n < 32; ++n) dxn[n] = dxc[n];
        }
        {   // it = 0: one input feature (the distance), nothing upstream: only the cotangents of the two layers' pre-activations
            const float4 *gz = reinterpret_cast<const float4 *>(a.dzw[0] + p * 32);
            float4 *gcm = reinterpret_cast<float4 *>(a.dzw_cm[0] + pc * 32), *gh = reinterpret_cast<float4 *>(a.dzh[0] + pc * 32);
            const float4 *x1 = reinterpret_cast<const float4 *>(a.px[1] + pc * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                gcm[q] = gz[q];
                const float4 t = x1[q];
                gh[q] = make_float4(dxn[4 * q] * (1.f - t.x * t.x), dxn[4 * q + 1] * (1.f - t.y * t.y), dxn[4 * q + 2] * (1.f - t.z * t.z),
                                    dxn[4 * q + 3] * (1.f - t.w * t.w));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ el-ion stream backward
// One warp per (walker, electron, ion): chain h^0 = [d, dx, dy, dz], h^{it+1} = res(tanh(h^it W + b), h^it);  conv_eI^it[i] = sum_J h^it[i, J] him^it[J].
//   cotangent of h^it from the convolution: dcei^it[b, i] * him^it[J];  pei^it[b, i, J] = h^it * dcei^it  (cotangent of him^it before the sum over i)
struct EionBwArgs {
    const float *r, *R;
    const float *w[DPE_MAX_ITER], *bias[DPE_MAX_ITER], *him[DPE_MAX_ITER], *dcei[DPE_MAX_ITER];
    float *ex[DPE_MAX_ITER], *dze[DPE_MAX_ITER], *pei[DPE_MAX_ITER];
    int dE[DPE_MAX_ITER];
    int n_iter, N, I;
    long n_rows;      // Bc N I
};

__global__ void __launch_bounds__(256) k_bw_eion(EionBwArgs a) {
    extern __shared__ float wsm[];                    // [it][32 x 33]
    for (int t = threadIdx.x; t < a.n_iter * WMAT; t += blockDim.x) wsm[t] = 0.f;
    __syncthreads();
    for (int it = 0; it + 1 < a.n_iter; ++it) stage_weight(wsm + it * WMAT, a.w[it], a.dE[it], a.dE[it + 1]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (long p = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5; p < a.n_rows; p += warps_total) {
    const long bi = p / a.I;
    const int J = (int)(p - bi * a.I);
    const float *ri = a.r + bi * 3;
    const float dx = ri[0] - a.R[J * 3], dy = ri[1] - a.R[J * 3 + 1], dz = ri[2] - a.R[J * 3 + 2];
    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
    float x[DPE_MAX_ITER], t[DPE_MAX_ITER];
    x[0] = lane == 0 ? d : (lane == 1 ? dx : (lane == 2 ? dy : (lane == 3 ? dz : 0.f)));
#pragma unroll
    for (int it = 0; it < DPE_MAX_ITER; ++it) {
        if (it < a.n_iter) {
            if (lane < a.dE[it]) a.ex[it][p * a.dE[it] + lane] = x[it];
            if (it + 1 < a.n_iter) {
                const int dn = a.dE[it + 1];
                const float z = warp_matvec(wsm + it * WMAT, a.dE[it], dn, x[it], lane) + (lane < dn ? a.bias[it][lane] : 0.f);
                t[it] = lane < dn ? tanhf(z) : 0.f;
                x[it + 1] = a.dE[it] == dn ? (x[it] + t[it]) * 0.70710678118654752f : t[it];
            }
        }
    }
    float dxn = 0.f;
#pragma unroll
    for (int it = DPE_MAX_ITER - 1; it >= 0; --it) {
        if (it < a.n_iter) {
            const int de = a.dE[it];
            const float dc = lane < de ? a.dcei[it][bi * de + lane] : 0.f;
            if (lane < de) a.pei[it][p * de + lane] = x[it] * dc;
            float dxc = lane < de ? dc * a.him[it][J * de + lane] : 0.f;
            if (it + 1 < a.n_iter) {
                const int dn = a.dE[it + 1];
                const bool res = de == dn;
                const float dzh = dxn * (res ? 0.70710678118654752f : 1.f) * (1.f - t[it] * t[it]);
                if (lane < dn) a.dze[it][p * dn + lane] = dzh;
                dxc += warp_matvec_t(wsm + it * WMAT, de, dn, dzh, lane);
                if (res) dxc += dxn * 0.70710678118654752f;
            }
            dxn = dxc;
        }
    }
    }
}

// dzhim[b, J, f] = (sum_i pei[b, i, J, f]) (1 - him[J, f]^2);   xion[b, J, :] = h_ion[Z_J] (the layer input, one row per (walker, ion))
__global__ void __launch_bounds__(256) k_bw_him(int N, int I, int dE, int F, const float *__restrict__ pei, const float *__restrict__ him,
                                                const float *__restrict__ emb_tab, const float *__restrict__ Zf, int z_min, float *__restrict__ dzhim,
                                                float *__restrict__ xion, float *__restrict__ onehot, int V) {
    const long b = blockIdx.x;
    for (int t = threadIdx.x; t < I * dE; t += blockDim.x) {
        const int J = t / dE, f = t - J * dE;
        float s = 0.f;
        for (int i = 0; i < N; ++i) s += pei[((b * N + i) * I + J) * (long)dE + f];
        const float h = him[J * dE + f];
        dzhim[(b * I + J) * (long)dE + f] = s * (1.f - h * h);
    }
    if (xion)
        for (int t = threadIdx.x; t < I * F; t += blockDim.x) {
            const int J = t / F, f = t - J * F;
            xion[(b * I + J) * (long)F + f] = emb_tab[(long)((int)Zf[J] - z_min) * F + f];
        }
    if (onehot)
        for (int t = threadIdx.x; t < I * V; t += blockDim.x) {
            const int J = t / V, v = t - J * V;
            onehot[(b * I + J) * (long)V + v] = ((int)Zf[J] - z_min) == v ? 1.f : 0.f;
        }
}

// ------------------------------------------------------------------------------------------------ workspace of one chunk
struct GradLayout {
    size_t x[DPE_MAX_ITER + 1], hm[DPE_MAX_ITER], mean[DPE_MAX_ITER], pw[DPE_MAX_ITER], ei[DPE_MAX_ITER];
    size_t add, bf, mo, det, ainv, coef, dbf, dy, dz, dx, sumdz, dmean, dzhm, dhmap, sumx;
    size_t dzw[DPE_MAX_ITER], dzw_cm[DPE_MAX_ITER], px[DPE_MAX_ITER], dzh[DPE_MAX_ITER], dcei[DPE_MAX_ITER], ex[DPE_MAX_ITER], dze[DPE_MAX_ITER], pei[DPE_MAX_ITER],
        dzhim[DPE_MAX_ITER];
    size_t xion, onehot, dhion, part, env_part, scr, tcs;
    size_t part_floats, tcs_floats, total;
    int ldx, env_splits, walkers_per_split;
};

static int d_one_in(const dpe_dims &d, int it) { return it == 0 ? 4 * d.n_ion : d.n_hidden_one_el[it - 1]; }
static int d_pair_in(const dpe_dims &d, int it) { return it == 0 ? 1 : d.n_hidden_two_el[it - 1]; }
static int d_eion_in(const dpe_dims &d, int it) { return it == 0 ? 4 : d.n_hidden_two_el[it - 1]; }

static void grad_plan(const dpe_dims &d, int Bc, GradLayout &L) {
    const int N = d.n_el, I = d.n_ion, nit = d.n_iterations, cols = d.n_dets * N, V = d.z_max - d.z_min + 1;
    int ldx = 0, max_din = 0, max_dout = 0, max_k = 0;
    for (int it = 0; it < nit; ++it) {
        const int din = d_one_in(d, it), km = din + d.emb_dim + d_eion_in(d, it);
        ldx = km > ldx ? km : ldx; max_k = ldx;
        max_din = din > max_din ? din : max_din;
        max_dout = d.n_hidden_one_el[it] > max_dout ? d.n_hidden_one_el[it] : max_dout;
    }
    if (max_dout > ldx) ldx = max_dout;
    ldx = (ldx + 3) & ~3;
    L.ldx = ldx;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off = gr_align(off + floats * sizeof(float)); return o; };
    const size_t R1 = (size_t)Bc * N, P2 = (size_t)Bc * N * N, R3 = R1 * I;
    for (int it = 0; it <= nit; ++it) L.x[it] = take(R1 * ldx);
    for (int it = 0; it < nit; ++it) {
        L.hm[it] = take(R1 * d.emb_dim);
        L.mean[it] = take((size_t)Bc * 2 * d_one_in(d, it));
        L.pw[it] = take(P2 * d.emb_dim);
        L.ei[it] = take(R1 * d_eion_in(d, it));
        L.dzw[it] = take(P2 * d.emb_dim);
        L.dzw_cm[it] = take(P2 * d.emb_dim);
        L.px[it] = take(P2 * d_pair_in(d, it));
        L.dzh[it] = take(it + 1 < nit ? P2 * d_pair_in(d, it + 1) : 0);
        L.dcei[it] = take(R1 * d_eion_in(d, it));
        L.ex[it] = take(R3 * d_eion_in(d, it));
        L.dze[it] = take(it + 1 < nit ? R3 * d_eion_in(d, it + 1) : 0);
        L.pei[it] = take(R3 * d_eion_in(d, it));
        L.dzhim[it] = take((size_t)Bc * I * d_eion_in(d, it));
    }
    L.add = take((size_t)Bc * max_dout);
    L.bf = take(d.use_taos ? R1 * I * cols : R1 * cols);       // TAO models: the per-ion orbital pre-factors g[b, i, J, col], overwritten by their cotangents
    L.mo = take(R1 * cols); L.dbf = take(d.use_taos ? 0 : R1 * cols);
    L.det = take((size_t)Bc * d.n_dets * 2);
    L.ainv = take((size_t)Bc * d.n_dets * N * N);
    L.coef = take((size_t)Bc * d.n_dets);
    L.dy = take(R1 * ldx); L.dz = take(R1 * max_dout); L.dx = take(R1 * ldx);
    L.sumdz = take((size_t)Bc * max_dout); L.dmean = take((size_t)Bc * 2 * max_din);
    L.dzhm = take(R1 * d.emb_dim); L.dhmap = take(R1 * max_din);
    L.sumx = take((size_t)Bc * (max_k + 1));
    L.xion = take((size_t)Bc * I * d.n_ion_features); L.onehot = take((size_t)Bc * I * V); L.dhion = take((size_t)Bc * I * d.n_ion_features);
    // scratch of the split products: the largest product is (k_main + 1) x (k_main + 1) or d_out^2 / cols^2, times up to 64 splits
    size_t biggest = (size_t)(max_k + 1) * (max_k + 1);
    if ((size_t)cols * cols > biggest) biggest = (size_t)cols * cols;
    if ((size_t)(max_k + 1) * 2 * max_din > biggest) biggest = (size_t)(max_k + 1) * 2 * max_din;
    L.part_floats = biggest * 64;
    L.part = take(L.part_floats);
    // transposed operands of the widest product on the tensor-core path: At [Mt][Rp] + Bt_hi, Bt_lo [Nb][Rp], rows padded to the K chunk
    {
        size_t wmax = (size_t)max_k + 1;
        if ((size_t)cols > wmax) wmax = cols;
        if ((size_t)max_dout > wmax) wmax = max_dout;
        L.tcs_floats = 3 * wmax * (R1 + 2048);
        L.tcs = take(L.tcs_floats);
    }
    {   // k_bw_orbitals: thread = orbital column, block row = walkers_per_split walkers; ~8 blocks of 64 threads per SM also for small batches
        const long col_blocks = (cols + 63) / 64;
        long wps = (long)Bc * col_blocks / (148L * 8);
        L.walkers_per_split = (int)std::max<long>(1, std::min<long>(64, wps));
    }
    L.env_splits = (Bc + L.walkers_per_split - 1) / L.walkers_per_split;
    L.env_part = take((size_t)L.env_splits * 2 * 2 * I * cols);
    L.scr = take((size_t)2 * (max_k + 1) * (max_k + 1) + (size_t)(max_k + 1) * 2 * max_din + (size_t)4 * max_din * max_din);
    L.total = off;
}

}  // namespace dpe

using namespace dpe;

// ------------------------------------------------------------------------------------------------ KFAC layer table
namespace dpe {

struct KfacLayer { char name[96]; int din, dout, has_bias, rows_per_walker; int64_t a_off, g_off; };

// The dense layers in the order of the canonical parameter leaves (include/dpe_b200.h): embedding lookup first, per iteration w_same, w_diff, h_map,
// h_ion_map, h_el, (h_same, h_diff, h_el_ion), then bf_up, bf_dn.
static int kfac_layers(const dpe_model *m, KfacLayer *out) {
    const dpe_dims &d = m->dims;
    const int N = d.n_el, U = d.n_up, D = N - U, I = d.n_ion;
    int n = 0;
    int64_t off = 0;
    auto add = [&](const char *fmt, int it, int din, int dout, int bias, int rpw) {
        KfacLayer &k = out[n++];
        snprintf(k.name, sizeof(k.name), fmt, it);
        k.din = din; k.dout = dout; k.has_bias = bias; k.rows_per_walker = rpw;
        k.a_off = off; off += (int64_t)(din + bias) * (din + bias);
        k.g_off = off; off += (int64_t)dout * dout;
    };
    add("wf/~/input/h_ion", 0, d.z_max - d.z_min + 1, d.n_ion_features, 0, I);
    for (int it = 0; it < d.n_iterations; ++it) {
        const int din = d_one_in(d, it), dP = d_pair_in(d, it), dE = d_eion_in(d, it);
        add("wf/fermi_net_embedding/symm_features_%d/convolutional_features/w_same/linear_0", it, dP, d.emb_dim, 1, U * U + D * D);
        add("wf/fermi_net_embedding/symm_features_%d/convolutional_features/w_diff/linear_0", it, dP, d.emb_dim, 1, 2 * U * D);
        add("wf/fermi_net_embedding/symm_features_%d/convolutional_features/h_map/linear_0", it, din, d.emb_dim, 1, N);
        add("wf/fermi_net_embedding/symm_features_%d/convolutional_features/h_ion_map/linear_0", it, d.n_ion_features, dE, 1, I);
        add("wf/fermi_net_embedding/h_el_%d/linear_0", it, 3 * din + d.emb_dim + dE, d.n_hidden_one_el[it], 1, N);
        if (it + 1 < d.n_iterations) {
            add("wf/fermi_net_embedding/h_same_%d/linear_0", it, dP, d.n_hidden_two_el[it], 1, U * U + D * D);
            add("wf/fermi_net_embedding/h_diff_%d/linear_0", it, dP, d.n_hidden_two_el[it], 1, 2 * U * D);
            add("wf/fermi_net_embedding/h_el_ion_%d/linear_0", it, dE, d.n_hidden_two_el[it], 1, N * I);
        }
    }
    if (d.use_taos) return n;          // the TAO head has no dense layer of its own (backflows come from the geometry cache)
    const int dl = d.n_hidden_one_el[d.n_iterations - 1], cols = d.n_dets * N;
    add("wf/~/orbitals/envelope_orbitals/bf_up/linear_0", 0, dl, cols, 0, U);
    add("wf/~/orbitals/envelope_orbitals/bf_dn/linear_0", 0, dl, cols, 0, D);
    return n;
}

// scatter the blocks of the h_el A factor into leaf order [h | mean_up mean_dn | conv | 1]
//   XX  [(km + 1) x (km + 1)]  from [x | 1]^T [x | 1] over (walker, electron) rows      (x = [h | conv])
//   XM  [(km + 1) x 2 d_in]    from [sum_i x | N]^T mean over walkers
//   MM  [2 d_in x 2 d_in]      from N mean^T mean over walkers
__global__ void k_assemble_hel(int d_in, int cw, const float *__restrict__ XX, const float *__restrict__ XM, const float *__restrict__ MM, float *__restrict__ A,
                               int accumulate) {
    const int km = d_in + cw, T = 3 * d_in + cw + 1;
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= (long)T * T) return;
    const int p = (int)(idx / T), q = (int)(idx - (long)p * T);
    // leaf position -> (kind, index): 0 x-part (index into [h | conv | 1]), 1 mean part
    auto map = [&](int t, int &kind, int &ix) {
        if (t < d_in) { kind = 0; ix = t; }
        else if (t < 3 * d_in) { kind = 1; ix = t - d_in; }
        else if (t < 3 * d_in + cw) { kind = 0; ix = d_in + (t - 3 * d_in); }
        else { kind = 0; ix = km; }
    };
    int kp, ip, kq, iq;
    map(p, kp, ip); map(q, kq, iq);
    float v;
    if (kp == 0 && kq == 0) v = XX[(long)ip * (km + 1) + iq];
    else if (kp == 0 && kq == 1) v = XM[(long)ip * 2 * d_in + iq];
    else if (kp == 1 && kq == 0) v = XM[(long)iq * 2 * d_in + ip];
    else v = MM[(long)ip * 2 * d_in + iq];
    A[idx] = (accumulate ? A[idx] : 0.f) + v;
}

// XX = [x | 1]^T x  ((km + 1) x km)  ->  the full symmetric (km + 1) x (km + 1) matrix [x | 1]^T [x | 1]  (corner = number of rows)
__global__ void k_symmetrize_xx(int km, const float *__restrict__ XX, float count, float *__restrict__ out) {
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const int T = km + 1;
    if (idx >= (long)T * T) return;
    const int p = (int)(idx / T), q = (int)(idx - (long)p * T);
    out[idx] = q < km ? XX[(long)p * km + q] : (p < km ? XX[(long)km * km + p] : count);
}

// (completion of the A factor of a biased layer whose [x | 1]^T x block ((din + 1) x din, row stride din + 1) has been accumulated: the ones column is the
// transpose of the ones row, the corner the row count -- k_kfac_complete below)
// The end of the KFAC pass for ALL layers in two launches (it was one k_complete_ones and two scaling launches per layer, ~90 launches of a few
// microseconds each): the table travels as a kernel argument.
constexpr int KFIN_MAX = 8 * DPE_MAX_ITER + 4;
struct KfacFinTab {
    int n;
    long long a_off[KFIN_MAX], g_off[KFIN_MAX];
    int din[KFIN_MAX], nA[KFIN_MAX], nG[KFIN_MAX], ones[KFIN_MAX];      // nA = (din + bias)^2, nG = dout^2; ones: complete the ones column first
    float rows[KFIN_MAX];
};
__global__ void __launch_bounds__(128) k_kfac_complete(float *__restrict__ kfac, const __grid_constant__ KfacFinTab t) {
    const int k = blockIdx.x;
    if (!t.ones[k]) return;
    const int din = t.din[k], T = din + 1;
    float *A = kfac + t.a_off[k];
    for (int p = threadIdx.x; p <= din; p += blockDim.x) A[(long)p * T + din] = p < din ? A[(long)din * T + p] : t.rows[k];
}
__global__ void __launch_bounds__(256) k_kfac_scale(float *__restrict__ kfac, const __grid_constant__ KfacFinTab t) {
    const int k = blockIdx.y >> 1, g = blockIdx.y & 1;
    float *p = kfac + (g ? t.g_off[k] : t.a_off[k]);
    const int n = g ? t.nG[k] : t.nA[k];
    const float s = 1.f / t.rows[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] *= s;
}

// ------------------------------------------------------------------------------------------------ one chunk: forward (kept) + backward
static int grad_chunk(dpe_model *m, const float *r, int Bc, const float *cot, float *grad, float *kfac, const KfacLayer *kl, bool accumulate, char *ws,
                      const GradLayout &L, float *logpsi2, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const int N = d.n_el, U = d.n_up, D = N - U, I = d.n_ion, nit = d.n_iterations, emb = d.emb_dim, cols = d.n_dets * N, F = d.n_ion_features;
    const int V = d.z_max - d.z_min + 1;
    const long R1 = (long)Bc * N, P2 = (long)Bc * N * N, R3 = R1 * I;
    const int ldx = L.ldx;
    auto fp = [&](size_t off) { return reinterpret_cast<float *>(ws + off); };
    int e;
    GradCtx g{m, s, fp(L.part), L.part_floats, cot, accumulate, fp(L.tcs), L.tcs_floats};
    auto leaf_off = [&](const float *p) { return (size_t)(p - m->params); };

    // ---------------- forward pass on the value channel, every intermediate kept
    if ((e = launch_features(m, r, Bc, 1, fp(L.x[0]), ldx, nullptr, s))) return e;
    {
        size_t ei_off[DPE_MAX_ITER], pw_off[DPE_MAX_ITER];
        for (int it = 0; it < nit; ++it) { ei_off[it] = (L.ei[it] - L.ei[0]) / sizeof(float); pw_off[it] = (L.pw[it] - L.pw[0]) / sizeof(float); }
        if ((e = launch_eion_stream(m, r, Bc, 1, fp(L.ei[0]), ei_off, s))) return e;
        if ((e = launch_pair_stream(m, r, Bc, 1, fp(L.pw[0]), pw_off, s))) return e;
    }
    for (int it = 0; it < nit; ++it) {
        const IterParams &p = m->it[it];
        float *x = fp(L.x[it]), *xn = fp(L.x[it + 1]), *hm = fp(L.hm[it]), *mean = fp(L.mean[it]), *add = fp(L.add);
        if ((e = dense_gemm(m, x, ldx, p.h_map.w, hm, emb, (int)R1, emb, p.d_in, s))) return e;
        if ((e = launch_act(m, hm, emb, (int)R1, 1, emb, p.h_map.b, nullptr, 1, s))) return e;
        if ((e = launch_conv(m, it, r, Bc, 1, hm, fp(L.pw[it]), fp(L.ei[it]), x, ldx, s))) return e;
        if ((e = launch_mean(m, x, ldx, Bc, 1, p.d_in, mean, s))) return e;
        if ((e = dense_gemm(m, mean, 2 * p.d_in, p.w_mean, add, p.d_out, Bc, p.d_out, 2 * p.d_in, s))) return e;
        if ((e = dense_gemm(m, x, ldx, p.w_main, xn, ldx, (int)R1, p.d_out, p.k_main, s))) return e;
        if ((e = launch_act(m, xn, ldx, (int)R1, 1, p.d_out, p.h_el.b, add, N, s))) return e;
    }
    const int dl = d.n_hidden_one_el[nit - 1];
    float *h_last = fp(L.x[nit]), *bf = fp(L.bf), *mo = fp(L.mo);
    if (d.use_taos) {
        // one projection against the cached backflow matrix for all electrons, then the exponential envelopes and the sum over ions
        if ((e = dense_gemm(m, h_last, ldx, m->tao_w, bf, I * cols, (int)R1, I * cols, dl, s))) return e;
        if ((e = launch_tao_orbitals(m, r, Bc, 1, bf, mo, s))) return e;
    } else {
        for (int sp = 0; sp < 2; ++sp)
            if ((e = dense_gemm_seg(m, h_last, ldx, m->bf_w[sp], bf, cols, Bc * (sp ? D : U), cols, dl, sp ? D : U, N, sp ? U : 0, s))) return e;
        DPE_CUDA(cudaMemcpyAsync(mo, bf, (size_t)R1 * cols * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if ((e = launch_envelope(m, r, Bc, 1, mo, s))) return e;
    }
    {
        const size_t sm_inv = ((size_t)N * (N + 1) + N) * sizeof(double) + (size_t)N * sizeof(int) + 8;
        static const bool inv_rows_off = getenv("DPE_DET_INV_ROWS") && atoi(getenv("DPE_DET_INV_ROWS")) == 0;
        const long n_mat = (long)Bc * d.n_dets;
        if (!inv_rows_off && N <= 64) {                   // one row per thread, in registers
            const unsigned g4 = (unsigned)((n_mat + 3) / 4);
            if (N <= 16) k_det_inverse_rows<16, 32><<<g4, 128, 0, s>>>(N, d.n_dets, n_mat, mo, fp(L.det), fp(L.ainv));
            else if (N <= 32) k_det_inverse_rows<32, 32><<<g4, 128, 0, s>>>(N, d.n_dets, n_mat, mo, fp(L.det), fp(L.ainv));
            else if (N <= 48) k_det_inverse_rows<48, 64><<<(unsigned)n_mat, 64, 0, s>>>(N, d.n_dets, n_mat, mo, fp(L.det), fp(L.ainv));
            else k_det_inverse_rows<64, 64><<<(unsigned)n_mat, 64, 0, s>>>(N, d.n_dets, n_mat, mo, fp(L.det), fp(L.ainv));
        } else if (N <= 32) k_det_inverse<32><<<Bc * d.n_dets, 32, sm_inv, s>>>(N, d.n_dets, mo, fp(L.det), fp(L.ainv));
        else k_det_inverse<64><<<Bc * d.n_dets, 64, sm_inv, s>>>(N, d.n_dets, mo, fp(L.det), fp(L.ainv));
    }
    DPE_LAUNCH_CHECK(m);
    k_bw_combine<<<(Bc * 32 + 127) / 128, 128, 0, s>>>(Bc, d.n_dets, fp(L.det), fp(L.coef), logpsi2);
    DPE_LAUNCH_CHECK(m);

    // ---------------- backward: orbitals
    int kbase[DPE_MAX_ITER], kidx_bf = 1;
    for (int it = 0; it < nit; ++it) { kbase[it] = kidx_bf; kidx_bf += (it + 1 < nit) ? 8 : 5; }
    float *dy = fp(L.dy);
    if (d.use_taos) {
        const long total = R1 * cols;
        k_bw_tao<<<(int)std::min<long>((total + 255) / 256, 148L * 32), 256, 0, s>>>(total, N, U, I, d.n_dets, r, m->R_dev, fp(L.coef), fp(L.ainv), m->tao_ex[0],
                                                                                   m->tao_ex[1], bf);
        DPE_LAUNCH_CHECK(m);
        // dh[b, i, :] = dg[b, i, :] @ W_tao^T
        if ((e = gemm_nt(m, bf, (long)I * cols, m->tao_w, (long)I * cols, dy, ldx, R1, dl, I * cols, false, s))) return e;
    } else {
    {
        dim3 grid((cols + 63) / 64, L.env_splits);
        if (I <= 4) k_bw_orbitals<4><<<grid, 64, 0, s>>>(Bc, N, U, I, d.n_dets, r, m->R_dev, fp(L.coef), fp(L.ainv), bf, m->sp_alpha[0], m->sp_alpha[1], m->alpha[0],
                                           m->alpha[1], m->env_w[0], m->env_w[1], cot, fp(L.dbf), fp(L.env_part), L.walkers_per_split);
        else if (I <= 8) k_bw_orbitals<8><<<grid, 64, 0, s>>>(Bc, N, U, I, d.n_dets, r, m->R_dev, fp(L.coef), fp(L.ainv), bf, m->sp_alpha[0], m->sp_alpha[1], m->alpha[0],
                                           m->alpha[1], m->env_w[0], m->env_w[1], cot, fp(L.dbf), fp(L.env_part), L.walkers_per_split);
        else k_bw_orbitals<16><<<grid, 64, 0, s>>>(Bc, N, U, I, d.n_dets, r, m->R_dev, fp(L.coef), fp(L.ainv), bf, m->sp_alpha[0], m->sp_alpha[1], m->alpha[0],
                                           m->alpha[1], m->env_w[0], m->env_w[1], cot, fp(L.dbf), fp(L.env_part), L.walkers_per_split);
        DPE_LAUNCH_CHECK(m);
        if (grad) {
            const long n = (long)I * cols;
            for (int sp = 0; sp < 2; ++sp) {
                k_bw_env_reduce<<<(int)((n + 31) / 32), 256, 0, s>>>(fp(L.env_part), L.env_splits, I, cols, sp, 0, grad + leaf_off(m->env_w[sp]), accumulate);
                DPE_LAUNCH_CHECK(m);
                k_bw_env_reduce<<<(int)((n + 31) / 32), 256, 0, s>>>(fp(L.env_part), L.env_splits, I, cols, sp, 1, grad + leaf_off(m->alpha[sp]), accumulate);
                DPE_LAUNCH_CHECK(m);
            }
        }
    }
    float *dbf = fp(L.dbf);
    for (int sp = 0; sp < 2; ++sp) {
        const int ns = sp ? D : U;
        const RowMap rm{ns, N, sp ? U : 0};             // the spin block of every walker
        const long rows = (long)Bc * ns;
        // dh[b, i, :] = dbf[b, i, :] @ W_bf^T
        if ((e = gemm_nt(m, dbf, cols, m->bf_w[sp], cols, dy, ldx, rows, dl, cols, false, s, rm))) return e;
        if (grad && (e = atb(g, grad + leaf_off(m->bf_w[sp]), cols, h_last, ldx, dl, false, dbf, cols, cols, rows, ns, cot, 1.f, -1, rm))) return e;
        if (kfac) {
            const KfacLayer &k = kl[kidx_bf + sp];
            if ((e = atb(g, kfac + k.a_off, dl, h_last, ldx, dl, false, h_last, ldx, dl, rows, ns, nullptr, 1.f, -1, rm))) return e;
            if ((e = atb(g, kfac + k.g_off, cols, dbf, cols, cols, false, dbf, cols, cols, rows, ns, nullptr, 0.5f, -1, rm))) return e;
        }
    }
    }

    // ---------------- backward: embedding iterations
    for (int it = nit - 1; it >= 0; --it) {
        const IterParams &p = m->it[it];
        float *x = fp(L.x[it]), *y = fp(L.x[it + 1]), *dz = fp(L.dz), *dx = fp(L.dx), *sumdz = fp(L.sumdz), *dmean = fp(L.dmean), *mean = fp(L.mean[it]);
        const int cw = emb + p.dE;
        k_bw_tanh<<<Bc, 256, 0, s>>>(N, p.d_out, dy, ldx, y, ldx, dz, sumdz);
        DPE_LAUNCH_CHECK(m);
        // dx = dz W_main^T  ([h | conv] columns),  dmean = sumdz W_mean^T
        if ((e = gemm_nt(m, dz, p.d_out, p.w_main, p.d_out, dx, ldx, R1, p.k_main, p.d_out, false, s))) return e;
        if ((e = gemm_nt(m, sumdz, p.d_out, p.w_mean, p.d_out, dmean, 2 * p.d_in, Bc, 2 * p.d_in, p.d_out, false, s))) return e;
        // h_el: gradient rows in leaf order h | mean_up mean_dn | conv | bias
        if (grad) {
            float *gw = grad + leaf_off(p.h_el.w);
            if ((e = atb(g, gw, p.d_out, x, ldx, p.d_in, false, dz, p.d_out, p.d_out, R1, N, cot, 1.f))) return e;
            if ((e = atb(g, gw + (size_t)p.d_in * p.d_out, p.d_out, mean, 2 * p.d_in, 2 * p.d_in, false, sumdz, p.d_out, p.d_out, Bc, 1, cot, 1.f))) return e;
            // conv rows and the bias row are contiguous in the leaf: [conv | 1]
            if ((e = atb(g, gw + (size_t)3 * p.d_in * p.d_out, p.d_out, x + p.d_in, ldx, cw, true, dz, p.d_out, p.d_out, R1, N, cot, 1.f))) return e;
        }
        if (kfac) {
            const KfacLayer &k = kl[kbase[it] + 4];
            float *sumx = fp(L.sumx);
            const int km = p.k_main;
            float *XX = fp(L.scr);                                      // [(km + 1) x km]
            float *XM = XX + (size_t)(km + 1) * km, *MM = XM + (size_t)(km + 1) * 2 * p.d_in;
            k_sum_electrons<<<Bc, 256, 0, s>>>(N, km, x, ldx, sumx);
            DPE_LAUNCH_CHECK(m);
            GradCtx g0 = g; g0.accumulate = false;
            if ((e = atb(g0, XX, km, x, ldx, km, true, x, ldx, km, R1, N, nullptr, 1.f))) return e;                 // [x | 1]^T x
            // the last column of [x | 1]^T [x | 1]: sums of x and the row count
            if ((e = atb(g0, XM, 2 * p.d_in, sumx, km + 1, km + 1, false, mean, 2 * p.d_in, 2 * p.d_in, Bc, 1, nullptr, 1.f))) return e;
            if ((e = atb(g0, MM, 2 * p.d_in, mean, 2 * p.d_in, 2 * p.d_in, false, mean, 2 * p.d_in, 2 * p.d_in, Bc, 1, nullptr, (float)N))) return e;
            // XX lacks its last COLUMN (x^T 1): symmetric, taken from the last row by the assembler -> build the full (km + 1)^2 matrix
            float *XXf = MM + (size_t)4 * p.d_in * p.d_in;
            k_symmetrize_xx<<<(int)(((long)(km + 1) * (km + 1) + 255) / 256), 256, 0, s>>>(km, XX, (float)R1, XXf);
            DPE_LAUNCH_CHECK(m);
            const long T = 3L * p.d_in + cw + 1;
            k_assemble_hel<<<(int)((T * T + 255) / 256), 256, 0, s>>>(p.d_in, cw, XXf, XM, MM, kfac + k.a_off, accumulate);
            DPE_LAUNCH_CHECK(m);
            if ((e = atb(g, kfac + k.g_off, p.d_out, dz, p.d_out, p.d_out, false, dz, p.d_out, p.d_out, R1, N, nullptr, 0.5f))) return e;
        }
        // SchNet convolution and h_map
        float *hm = fp(L.hm[it]), *dzhm = fp(L.dzhm), *dhmap = fp(L.dhmap);
        k_bw_conv<<<Bc, 256, 0, s>>>(N, emb, dx, ldx, p.d_in, hm, fp(L.pw[it]), fp(L.dzw[it]), dzhm);
        DPE_LAUNCH_CHECK(m);
        if ((e = gemm_nt(m, dzhm, emb, p.h_map.w, emb, dhmap, p.d_in, R1, p.d_in, emb, false, s))) return e;
        if (grad && (e = atb(g, grad + leaf_off(p.h_map.w), emb, x, ldx, p.d_in, true, dzhm, emb, emb, R1, N, cot, 1.f))) return e;
        if (kfac) {
            const KfacLayer &k = kl[kbase[it] + 2];
            if ((e = atb(g, kfac + k.a_off, p.d_in + 1, x, ldx, p.d_in, true, x, ldx, p.d_in, R1, N, nullptr, 1.f))) return e;
            // the ones column of A: the symmetric completion is done at the end (finish_kfac)
            if ((e = atb(g, kfac + k.g_off, emb, dzhm, emb, emb, false, dzhm, emb, emb, R1, N, nullptr, 0.5f))) return e;
        }
        k_bw_gather<<<(unsigned)R1, 128, 0, s>>>(N, U, p.d_in, emb, p.dE, dx, ldx, dhmap, dmean, it > 0 ? dy : nullptr, ldx, fp(L.dcei[it]));
        DPE_LAUNCH_CHECK(m);
    }

    // ---------------- backward: pair stream and el-ion stream (recurrences over the iterations)
    {
        PairBwArgs a;
        a.r = r; a.n_iter = nit; a.N = N; a.U = U; a.emb = emb; a.n_pairs = P2;
        for (int it = 0; it < nit; ++it) {
            const IterParams &p = m->it[it];
            a.dP[it] = p.dP;
            a.ww[it][0] = p.w_same.w; a.wb[it][0] = p.w_same.b; a.ww[it][1] = p.w_diff.w; a.wb[it][1] = p.w_diff.b;
            a.hw[it][0] = p.h_same.w; a.hb[it][0] = p.h_same.b; a.hw[it][1] = p.h_diff.w; a.hb[it][1] = p.h_diff.b;
            a.dzw[it] = fp(L.dzw[it]); a.dzw_cm[it] = fp(L.dzw_cm[it]); a.px[it] = fp(L.px[it]); a.dzh[it] = fp(L.dzh[it]);
        }
        const size_t sm_pair = (size_t)nit * 4 * WMAT * sizeof(float), sm_eion = (size_t)nit * WMAT * sizeof(float);
        if ((e = opt_in_smem(m, KID_BW_PAIR, k_bw_pair))) return e;       // per model / device, as every other kernel with more than 48 KB
        if ((e = opt_in_smem(m, KID_BW_EION, k_bw_eion))) return e;
        static const bool rows_off = getenv("DPE_BW_PAIR_ROWS") && atoi(getenv("DPE_BW_PAIR_ROWS")) == 0;
        bool rows_ok = !rows_off && emb == 32 && nit >= 2 && a.dP[0] == 1;
        for (int it = 1; it < nit; ++it) rows_ok = rows_ok && a.dP[it] == 32;
        if (rows_ok) {                // default widths: one thread per pair, operands in registers
            const size_t sm_rows = ((size_t)nit * 4 * 1024 + (size_t)nit * 64) * sizeof(float);
            if ((e = opt_in_smem(m, KID_BW_PAIR_ROWS, k_bw_pair_rows))) return e;
            const int per_sm = 1;
            k_bw_pair_rows<<<(unsigned)std::min<long>((P2 + PBR_THREADS - 1) / PBR_THREADS, 148L * per_sm), PBR_THREADS, sm_rows, s>>>(a);
        } else {
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200u << 10) / sm_pair));
            k_bw_pair<<<(unsigned)std::min<long>((P2 * 32 + 255) / 256, 148L * per_sm), 256, sm_pair, s>>>(a);
        }
        DPE_LAUNCH_CHECK(m);
        EionBwArgs b;
        b.r = r; b.R = m->R_dev; b.n_iter = nit; b.N = N; b.I = I; b.n_rows = R3;
        for (int it = 0; it < nit; ++it) {
            const IterParams &p = m->it[it];
            b.dE[it] = p.dE; b.w[it] = p.h_el_ion.w; b.bias[it] = p.h_el_ion.b; b.him[it] = p.him; b.dcei[it] = fp(L.dcei[it]);
            b.ex[it] = fp(L.ex[it]); b.dze[it] = fp(L.dze[it]); b.pei[it] = fp(L.pei[it]);
        }
        k_bw_eion<<<(unsigned)std::min<long>((R3 * 32 + 255) / 256, 148L * 8), 256, sm_eion, s>>>(b);
        DPE_LAUNCH_CHECK(m);
    }
    float *dhion = fp(L.dhion);
    for (int it = 0; it < nit; ++it) {
        const IterParams &p = m->it[it];
        for (int sd = 0; sd < 2; ++sd) {
            const Dense &wl = sd ? p.w_diff : p.w_same, &hl = sd ? p.h_diff : p.h_same;
            // class-major rows (k_bw_pair): the rows of this spin class are one contiguous range; rows per walker = pairs of the class
            const int rpw = sd ? 2 * U * D : U * U + D * D;
            const long row0 = sd ? (long)Bc * (U * U + D * D) : 0, rows_c = (long)Bc * rpw;
            if (rows_c == 0) continue;
            const float *px = fp(L.px[it]) + row0 * p.dP, *dzw = fp(L.dzw_cm[it]) + row0 * emb;
            if (grad && (e = atb(g, grad + leaf_off(wl.w), emb, px, p.dP, p.dP, true, dzw, emb, emb, rows_c, rpw, cot, 1.f))) return e;
            if (kfac) {
                const KfacLayer &k = kl[kbase[it] + sd];
                if ((e = atb(g, kfac + k.a_off, p.dP + 1, px, p.dP, p.dP, true, px, p.dP, p.dP, rows_c, rpw, nullptr, 1.f))) return e;
                if ((e = atb(g, kfac + k.g_off, emb, dzw, emb, emb, false, dzw, emb, emb, rows_c, rpw, nullptr, 0.5f))) return e;
            }
            if (it + 1 < nit) {
                const int dn = m->it[it + 1].dP;
                const float *dzh = fp(L.dzh[it]) + row0 * dn;
                if (grad && (e = atb(g, grad + leaf_off(hl.w), dn, px, p.dP, p.dP, true, dzh, dn, dn, rows_c, rpw, cot, 1.f))) return e;
                if (kfac) {
                    const KfacLayer &k = kl[kbase[it] + 5 + sd];
                    // the A factor equals that of w_same / w_diff (same input rows): copied at the end (dpe_param_gradient)
                    if ((e = atb(g, kfac + k.g_off, dn, dzh, dn, dn, false, dzh, dn, dn, rows_c, rpw, nullptr, 0.5f))) return e;
                }
            }
        }
        if (it + 1 < nit) {
            const int dn = m->it[it + 1].dE;
            if (grad && (e = atb(g, grad + leaf_off(p.h_el_ion.w), dn, fp(L.ex[it]), p.dE, p.dE, true, fp(L.dze[it]), dn, dn, R3, N * I, cot, 1.f))) return e;
            if (kfac) {
                const KfacLayer &k = kl[kbase[it] + 7];
                if ((e = atb(g, kfac + k.a_off, p.dE + 1, fp(L.ex[it]), p.dE, p.dE, true, fp(L.ex[it]), p.dE, p.dE, R3, N * I, nullptr, 1.f))) return e;
                if ((e = atb(g, kfac + k.g_off, dn, fp(L.dze[it]), dn, dn, false, fp(L.dze[it]), dn, dn, R3, N * I, nullptr, 0.5f))) return e;
            }
        }
        // ion-level layers: h_ion_map (rows = (walker, ion)) and the embedding lookup
        float *dzhim = fp(L.dzhim[it]), *xion = fp(L.xion), *onehot = fp(L.onehot);
        k_bw_him<<<Bc, 256, 0, s>>>(N, I, p.dE, F, fp(L.pei[it]), p.him, m->h_ion_emb, m->Z_dev, d.z_min, dzhim, it == 0 ? xion : nullptr,
                                    it == 0 ? onehot : nullptr, V);
        DPE_LAUNCH_CHECK(m);
        const long RI = (long)Bc * I;
        if (grad && (e = atb(g, grad + leaf_off(p.h_ion_map.w), p.dE, xion, F, F, true, dzhim, p.dE, p.dE, RI, I, cot, 1.f))) return e;
        if (kfac) {
            const KfacLayer &k = kl[kbase[it] + 3];
            if ((e = atb(g, kfac + k.a_off, F + 1, xion, F, F, true, xion, F, F, RI, I, nullptr, 1.f))) return e;
            if ((e = atb(g, kfac + k.g_off, p.dE, dzhim, p.dE, p.dE, false, dzhim, p.dE, p.dE, RI, I, nullptr, 0.5f))) return e;
        }
        if ((e = gemm_nt(m, dzhim, p.dE, p.h_ion_map.w, p.dE, dhion, F, RI, F, p.dE, it > 0, s))) return e;
    }
    {
        const long RI = (long)Bc * I;
        if (grad && (e = atb(g, grad + leaf_off(m->h_ion_emb), F, fp(L.onehot), V, V, false, dhion, F, F, RI, I, cot, 1.f))) return e;
        if (kfac) {
            const KfacLayer &k = kl[0];
            if ((e = atb(g, kfac + k.a_off, V, fp(L.onehot), V, V, false, fp(L.onehot), V, V, RI, I, nullptr, 1.f))) return e;
            if ((e = atb(g, kfac + k.g_off, F, dhion, F, F, false, dhion, F, F, RI, I, nullptr, 0.5f))) return e;
        }
    }
    return DPE_OK;
}

}  // namespace dpe

extern "C" {

int32_t dpe_kfac_layer_count(const dpe_model *m) {
    if (!m) return 0;
    KfacLayer kl[8 * DPE_MAX_ITER + 4];
    return kfac_layers(m, kl);
}

int64_t dpe_kfac_floats(const dpe_model *m) {
    if (!m) return 0;
    KfacLayer kl[8 * DPE_MAX_ITER + 4];
    const int n = kfac_layers(m, kl);
    return kl[n - 1].g_off + (int64_t)kl[n - 1].dout * kl[n - 1].dout;
}

int dpe_kfac_layer(const dpe_model *m, int32_t index, char *name, int32_t name_len, int32_t *din, int32_t *dout, int32_t *has_bias, int32_t *rows_per_walker,
                   int64_t *a_offset, int64_t *g_offset) {
    if (!m) return set_error(DPE_ERR_ARG, "kfac_layer: null model");
    KfacLayer kl[8 * DPE_MAX_ITER + 4];
    const int n = kfac_layers(m, kl);
    if (index < 0 || index >= n) return set_error(DPE_ERR_ARG, "kfac_layer: index %d outside [0, %d)", index, n);
    const KfacLayer &k = kl[index];
    if (name && name_len > 0) { strncpy(name, k.name, name_len - 1); name[name_len - 1] = 0; }
    if (din) *din = k.din;
    if (dout) *dout = k.dout;
    if (has_bias) *has_bias = k.has_bias;
    if (rows_per_walker) *rows_per_walker = k.rows_per_walker;
    if (a_offset) *a_offset = k.a_off;
    if (g_offset) *g_offset = k.g_off;
    return DPE_OK;
}

size_t dpe_gradient_workspace_bytes(const dpe_model *m, int32_t n_walkers) {
    if (!m || n_walkers <= 0) return 0;
    GradLayout L;
    grad_plan(m->dims, n_walkers, L);
    return L.total + 256;
}

int dpe_param_gradient(dpe_model *m, const float *r_dev, int32_t n_walkers, const float *cotangent_dev, float *grad_dev, float *kfac_dev,
                       float *log_psi_sqr_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!m || !r_dev || !workspace_dev || n_walkers <= 0 || (!grad_dev && !kfac_dev)) return set_error(DPE_ERR_ARG, "param_gradient: bad argument");
    if (grad_dev && !cotangent_dev) return set_error(DPE_ERR_ARG, "param_gradient: the gradient needs the per-walker cotangents");
    if (m->dims.use_taos && !m->tao_set) return set_error(DPE_ERR_STATE, "param_gradient: set_tao_cache must be called first");
    if (!m->params_set || !m->geom_set) return set_error(DPE_ERR_STATE, "set_params and set_geometry must be called first");
    const dpe_dims &d = m->dims;
    if (d.emb_dim > 32 || d.n_ion_features > 32) return set_error(DPE_ERR_UNSUPPORTED, "param_gradient: pair / ion layers wider than 32");
    for (int it = 0; it + 1 < d.n_iterations; ++it)
        if (d.n_hidden_two_el[it] > 32) return set_error(DPE_ERR_UNSUPPORTED, "param_gradient: pair / ion layers wider than 32");
    cudaStream_t s = (cudaStream_t)stream;
    GradLayout L;
    int chunk = n_walkers;
    grad_plan(d, chunk, L);
    if (L.total > workspace_bytes) {                 // largest chunk that fits
        int lo = 0, hi = n_walkers;
        while (hi - lo > 1) {
            const int mid = lo + (hi - lo) / 2;
            grad_plan(d, mid, L);
            if (L.total <= workspace_bytes) lo = mid; else hi = mid;
        }
        chunk = lo;
    }
    if (chunk < 1) return set_error(DPE_ERR_WORKSPACE, "gradient workspace of %zu bytes cannot hold one walker", workspace_bytes);
    KfacLayer kl[8 * DPE_MAX_ITER + 4];
    memset(kl, 0, sizeof(kl));
    const int n_layers = kfac_layers(m, kl);
    for (int off = 0; off < n_walkers; off += chunk) {
        const int Bc = n_walkers - off < chunk ? n_walkers - off : chunk;
        grad_plan(d, Bc, L);
        int e = grad_chunk(m, r_dev + (size_t)off * d.n_el * 3, Bc, cotangent_dev ? cotangent_dev + off : nullptr, grad_dev, kfac_dev, kl, off > 0,
                           (char *)workspace_dev, L, log_psi_sqr_dev ? log_psi_sqr_dev + off : nullptr, s);
        if (e) return e;
    }
    if (kfac_dev) {
        for (int it = 0, base = 1; it < d.n_iterations; ++it) {           // h_same / h_diff see the rows w_same / w_diff see: one A factor serves both
            if (it + 1 < d.n_iterations)
                for (int sd = 0; sd < 2; ++sd) {
                    const KfacLayer &src = kl[base + sd], &dst = kl[base + 5 + sd];
                    if (int e = check_cuda(cudaMemcpyAsync(kfac_dev + dst.a_off, kfac_dev + src.a_off, sizeof(float) * (src.din + 1) * (src.din + 1), cudaMemcpyDeviceToDevice, s), __func__)) return e;
                }
            base += it + 1 < d.n_iterations ? 8 : 5;
        }
        KfacFinTab tab;
        tab.n = n_layers;
        int n_max = 1;
        for (int k = 0; k < n_layers; ++k) {
            const KfacLayer &l = kl[k];
            const bool is_hel = strstr(l.name, "/h_el_") != nullptr && strstr(l.name, "/h_el_ion_") == nullptr;          // assembled with its ones row and column already
            tab.a_off[k] = l.a_off; tab.g_off[k] = l.g_off; tab.din[k] = l.din;
            tab.nA[k] = (l.din + l.has_bias) * (l.din + l.has_bias); tab.nG[k] = l.dout * l.dout;
            tab.ones[k] = (l.has_bias && !is_hel) ? 1 : 0;
            tab.rows[k] = (float)((double)n_walkers * l.rows_per_walker);
            n_max = std::max(n_max, std::max(tab.nA[k], tab.nG[k]));
        }
        k_kfac_complete<<<n_layers, 128, 0, s>>>(kfac_dev, tab);
        DPE_LAUNCH_CHECK(m);
        k_kfac_scale<<<dim3((unsigned)std::min(32, (n_max + 2047) / 2048), (unsigned)(2 * n_layers)), 256, 0, s>>>(kfac_dev, tab);
        DPE_LAUNCH_CHECK(m);
    }
    return DPE_OK;
}

}  // extern "C"
