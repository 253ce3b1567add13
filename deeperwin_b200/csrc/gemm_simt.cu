// FP32 SIMT GEMM: C[m, n] = sum_k A[m, k] W[k, n] with segmented row addressing for A and C.
// Baseline dense-layer path (true FP32, as the reference demands: process_molecule.py:20-29) and
// the fallback for shapes the tcgen05 3xTF32 kernel (gemm_tc.cu) does not take.
// Classic register-tiled kernel: BK=16 k-slabs double-buffered through shared memory, A stored
// transposed so that the inner loop is LDS.128 + 64 (or 32) FFMA per k.
#include "dpe_internal.cuh"

namespace dpe {

constexpr int BK = 16;

__device__ __forceinline__ long seg_row(int m, int seg_len, int seg_stride, int seg_off) {
    int q = m / seg_len;
    return (long)q * seg_stride + seg_off + (m - q * seg_len);
}

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) k_gemm_simt(GemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int SA = BM + 4;  // padded stride (floats), keeps float4 alignment
    constexpr int SB = BN + 4;
    __shared__ __align__(16) float As[2][BK][SA];
    __shared__ __align__(16) float Bs[2][BK][SB];

    const int tid = threadIdx.x;
    const int n_tiles_n = (g.N + BN - 1) / BN;   // 1-D grid, n-tiles of one m-tile adjacent (share A in L2)
    const int n0 = (blockIdx.x % n_tiles_n) * BN;
    const int m0 = (blockIdx.x / n_tiles_n) * BM;
    constexpr int TXN = BN / TN;           // threads along n
    const int tx = tid % TXN, ty = tid / TXN;

    // --- global->register staging of one k-slab
    constexpr int A_F4 = BM * BK / 4;      // float4 per A slab
    constexpr int A_PER = (A_F4 + NT - 1) / NT;
    constexpr int B_EL = BK * BN;
    constexpr int B_PER = (B_EL + NT - 1) / NT;
    float4 a_reg[A_PER];
    float b_reg[B_PER];
    const float *a_ptr[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int idx = tid + i * NT;            // float4 index: row = idx / (BK/4), kq = idx % (BK/4)
        int row = idx / (BK / 4);
        int m = m0 + row;
        a_ok[i] = (idx < A_F4) && (m < g.M);
        a_ptr[i] = a_ok[i] ? g.A + seg_row(m, g.a_seg_len, g.a_seg_stride, g.a_seg_off) * g.lda + (idx % (BK / 4)) * 4 : g.A;
    }

    auto load_slab = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int kq = ((tid + i * NT) % (BK / 4)) * 4;
            if (a_ok[i] && k0 + kq < g.K)   // K % 4 == 0 is guaranteed by the caller
                a_reg[i] = *reinterpret_cast<const float4 *>(a_ptr[i] + k0);
            else
                a_reg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int idx = tid + i * NT;
            int k = idx / BN, n = idx % BN;
            b_reg[i] = (idx < B_EL && k0 + k < g.K && n0 + n < g.N) ? g.W[(long)(k0 + k) * g.ldw + n0 + n] : 0.f;
        }
    };
    auto store_slab = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int idx = tid + i * NT;
            if (idx < A_F4) {
                int row = idx / (BK / 4), kq = (idx % (BK / 4)) * 4;
                As[buf][kq + 0][row] = a_reg[i].x;
                As[buf][kq + 1][row] = a_reg[i].y;
                As[buf][kq + 2][row] = a_reg[i].z;
                As[buf][kq + 3][row] = a_reg[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int idx = tid + i * NT;
            if (idx < B_EL) Bs[buf][idx / BN][idx % BN] = b_reg[i];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (g.K + BK - 1) / BK;
    load_slab(0);
    store_slab(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_slab((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            // rows: TM/4 groups of 4 spaced BM/(TM/4) apart -> LDS.128, broadcast within the warp
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                float4 v = *reinterpret_cast<const float4 *>(&As[buf][k][q * (BM / (TM / 4)) + ty * 4]);
                a[q * 4 + 0] = v.x; a[q * 4 + 1] = v.y; a[q * 4 + 2] = v.z; a[q * 4 + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < TN / 4; ++q) {
                float4 v = *reinterpret_cast<const float4 *>(&Bs[buf][k][q * (BN / (TN / 4)) + tx * 4]);
                b[q * 4 + 0] = v.x; b[q * 4 + 1] = v.y; b[q * 4 + 2] = v.z; b[q * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_slab(buf ^ 1);
            __syncthreads();
        }
    }

    // --- epilogue
    const bool vec_ok = ((g.ldc & 3) == 0) && ((g.c_col_off & 3) == 0) && ((reinterpret_cast<size_t>(g.C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int row = (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4);
        int m = m0 + row;
        if (m >= g.M) continue;
        float *crow = g.C + seg_row(m, g.c_seg_len, g.c_seg_stride, g.c_seg_off) * g.ldc + g.c_col_off;
#pragma unroll
        for (int q = 0; q < TN / 4; ++q) {
            int n = n0 + q * (BN / (TN / 4)) + tx * 4;
            if (vec_ok && n + 3 < g.N) {
                *reinterpret_cast<float4 *>(crow + n) = make_float4(acc[i][q * 4], acc[i][q * 4 + 1], acc[i][q * 4 + 2], acc[i][q * 4 + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < g.N) crow[n + j] = acc[i][q * 4 + j];
            }
        }
    }
}

int launch_gemm_simt(dpe_model *m, const GemmArgs &g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return DPE_OK;
    if ((g.K & 3) || (g.lda & 3) || (reinterpret_cast<size_t>(g.A) & 15))
        return set_error(DPE_ERR_UNSUPPORTED, "gemm: K=%d / lda=%d must be multiples of 4 and A 16B aligned", g.K, g.lda);
    if (g.N > 64) {
        dim3 grid((unsigned)(((g.N + 127) / 128) * (long)((g.M + 127) / 128)));
        k_gemm_simt<128, 128, 8, 8><<<grid, 256, 0, s>>>(g);
        m->last_gemm_class = 0;
    } else if (g.N > 32) {
        dim3 grid((unsigned)(((g.N + 63) / 64) * (long)((g.M + 127) / 128)));
        k_gemm_simt<128, 64, 8, 4><<<grid, 256, 0, s>>>(g);
        m->last_gemm_class = 1;
    } else {
        dim3 grid((unsigned)(((g.N + 31) / 32) * (long)((g.M + 255) / 256)));
        k_gemm_simt<256, 32, 8, 4><<<grid, 256, 0, s>>>(g);
        m->last_gemm_class = 2;
    }
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

}  // namespace dpe
