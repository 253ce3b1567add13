// Pair stream of the forward pass on the tensor cores.
// Reference: model/embeddings/ferminet_embedding.py:45-104, 190-267 (w_same / w_diff, h_same / h_diff: tanh(x W + b) with residual on the
// pair features, which depend on the scalar distance |r_i - r_j| only).
//
// The SIMT kernel (streams.cu: k_pair_stream_fwd) runs at ~60 % of the FP32 CUDA-core peak and is the largest kernel of a Metropolis step.
// Here one thread owns one (walker, pair i <= j) row for the whole chain of iterations; the 32 -> (32 | 32) dense layers of iterations >= 1
// are tcgen05 MMAs  D[128 rows, 64] = X[128, 32] . [Ww | Wh]  (3xTF32, FP32 accumulate in TMEM):
//   * the tf32-split weights of all iterations and both spin classes stay resident in shared memory (K-major SWIZZLE_64B slabs, written once per CTA),
//   * the thread writes its row of X (hi / lo) into the operand tile, one elected thread issues 12 MMAs of M128 x N64 x K8, the thread reads its
//     64 outputs back with tcgen05.ld, applies bias + tanh (+ residual), stores w and keeps x in registers for the next iteration,
//   * three 128-row tiles are in flight per CTA (own operand buffers, own 64 TMEM columns, own mbarriers) so the MMA latency of one hides
//     under the tanh work of the others.
// Rows are ordered [spin class][walker][pair]: a tile is class-homogeneous.  Iteration 0 (one input feature, the distance) is an outer product.
#include <cstdlib>
#include "dpe_internal.cuh"
#include "tc_common.cuh"

namespace dpe {

constexpr int PT_GROUPS = 3;
constexpr int PT_THREADS = 32 + PT_GROUPS * 128;
constexpr int PT_W_BYTES = 2 * 64 * TC_ROWB;        // one of hi / lo of a [64 x 32] weight: two K slabs of 64 rows x 64 B
constexpr int PT_A_BYTES = 2 * 128 * TC_ROWB;       // one of hi / lo of a [128 x 32] operand tile

struct PairTcArgs {
    const float *r;
    const float *ww[DPE_MAX_ITER][2], *wb[DPE_MAX_ITER][2], *hw[DPE_MAX_ITER][2], *hb[DPE_MAX_ITER][2];
    float *out[DPE_MAX_ITER];
    int n_iter, N, U, n_walkers;
    int n_cls[2];          // pairs (i <= j) per walker: same spin, different spin
    long tiles[2];         // 128-row tiles per class
    float corr;            // accumulation-bias compensation of the K = 32 products
};

__device__ __forceinline__ uint32_t sw64_offset(int row, int k16) {        // byte offset of element (row, k16) inside a K slab (64-byte rows, SWIZZLE_64B)
    return (uint32_t)(row * TC_ROWB + ((((k16 >> 2) ^ ((row >> 1) & 3)) << 4) | ((k16 & 3) << 2)));
}

__global__ void __launch_bounds__(PT_THREADS, 1) k_pair_stream_tc(PairTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n_l = a.n_iter - 1;                                         // tensor-core layers (iterations 1 .. n_iter - 1)
    uint8_t *w_base = smem;                                                // [layer][class][hi | lo] PT_W_BYTES each
    uint8_t *a_base = w_base + (size_t)n_l * 2 * 2 * PT_W_BYTES;           // [group][hi | lo] PT_A_BYTES each
    float *tab = reinterpret_cast<float *>(a_base + (size_t)PT_GROUPS * 2 * PT_A_BYTES);
    float *bias_w = tab;                                                   // [iter][class][32]
    float *bias_h = bias_w + a.n_iter * 64;                                // [iter][class][32]
    float *w0 = bias_h + a.n_iter * 64;                                    // [class][32]  iteration-0 weights (one input feature)
    float *h0 = w0 + 64;                                                   // [class][32]
    int *pair_tab = reinterpret_cast<int *>(h0 + 64);                      // same-spin pairs, then different-spin pairs; (i << 8) | j
    uint64_t *bars = reinterpret_cast<uint64_t *>(pair_tab + ((a.n_cls[0] + a.n_cls[1] + 1) & ~1));
    uint64_t *bar_a = bars, *bar_d = bars + PT_GROUPS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_d + PT_GROUPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int g = 0; g < PT_GROUPS; ++g) { mbar_init(&bar_a[g], 128); mbar_init(&bar_d[g], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int ps = 0, pd = a.n_cls[0];
        for (int i = 0; i < a.N; ++i)
            for (int j = i; j < a.N; ++j) {
                if ((i < a.U) == (j < a.U)) pair_tab[ps++] = (i << 8) | j;
                else pair_tab[pd++] = (i << 8) | j;
            }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // resident weights: Wt[n][k] = [Ww | Wh][k][n], tf32 hi / lo, K-major SWIZZLE_64B slabs
    for (int t = threadIdx.x; t < n_l * 2 * 64 * 32; t += blockDim.x) {
        const int k = t & 31, n = (t >> 5) & 63, c = (t >> 11) & 1, l = t >> 12, it = l + 1;
        float w = 0.f;
        if (n < 32) w = a.ww[it][c][k * 32 + n];
        else if (it + 1 < a.n_iter) w = a.hw[it][c][k * 32 + n - 32];
        const float hi = rna_tf32(w), lo = rna_tf32(w - hi);
        uint8_t *dst = w_base + (size_t)(l * 2 + c) * 2 * PT_W_BYTES + (k >> 4) * (64 * TC_ROWB) + sw64_offset(n, k & 15);
        *reinterpret_cast<float *>(dst) = hi;
        *reinterpret_cast<float *>(dst + PT_W_BYTES) = lo;
    }
    for (int t = threadIdx.x; t < a.n_iter * 64; t += blockDim.x) {
        const int f = t & 31, c = (t >> 5) & 1, it = t >> 6;
        bias_w[t] = a.wb[it][c][f];
        bias_h[t] = it + 1 < a.n_iter ? a.hb[it][c][f] : 0.f;
    }
    for (int t = threadIdx.x; t < 64; t += blockDim.x) {
        const int f = t & 31, c = t >> 5;
        w0[t] = a.ww[0][c][f];
        h0[t] = a.n_iter > 1 ? a.hw[0][c][f] : 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long n_tiles = a.tiles[0] + a.tiles[1];
    const long stride = (long)gridDim.x * PT_GROUPS;

    if (warp == 0) {
        if (lane == 0 && n_l > 0) {
            const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
            uint32_t ph[PT_GROUPS] = {0, 0, 0};
            for (long u0 = (long)blockIdx.x * PT_GROUPS; u0 < n_tiles; u0 += stride)
                for (int l = 0; l < n_l; ++l)
                    for (int g = 0; g < PT_GROUPS; ++g) {
                        const long u = u0 + g;
                        if (u >= n_tiles) continue;
                        const int c = u >= a.tiles[0] ? 1 : 0;
                        mbar_wait(&bar_a[g], ph[g]);
                        ph[g] ^= 1;
                        tc_fence_after();
                        const uint32_t d = tmem_base + g * 64;
                        const uint32_t xh = smem_u32(a_base + (size_t)g * 2 * PT_A_BYTES), xl = xh + PT_A_BYTES;
                        const uint32_t wh = smem_u32(w_base + (size_t)(l * 2 + c) * 2 * PT_W_BYTES), wl = wh + PT_W_BYTES;
                        const uint32_t idesc = l + 2 < a.n_iter ? idesc64 : idesc32;      // the last iteration has no h layer
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int kk = 0; kk < TC_BK / 8; ++kk) {
                                const uint32_t ko = kk * 32;
                                const uint64_t dxh = make_desc_sw64(xh + kb * (128 * TC_ROWB) + ko), dxl = make_desc_sw64(xl + kb * (128 * TC_ROWB) + ko);
                                const uint64_t dwh = make_desc_sw64(wh + kb * (64 * TC_ROWB) + ko), dwl = make_desc_sw64(wl + kb * (64 * TC_ROWB) + ko);
                                tc_mma_tf32(d, dxh, dwl, idesc, (kb | kk) ? 1u : 0u);
                                tc_mma_tf32(d, dxl, dwh, idesc, 1u);
                                tc_mma_tf32(d, dxh, dwh, idesc, 1u);
                            }
                        tc_commit(&bar_d[g]);
                    }
        }
    } else {
        const int g = (warp - 1) >> 2;
        const int row = (warp & 3) * 32 + lane;            // TMEM lane quarter of a warp is fixed by warp id % 4
        uint8_t *ah = a_base + (size_t)g * 2 * PT_A_BYTES, *al = ah + PT_A_BYTES;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + g * 64;
        uint32_t ph = 0;
        const int N = a.N;
        for (long u = (long)blockIdx.x * PT_GROUPS + g; u < n_tiles; u += stride) {
            const int c = u >= a.tiles[0] ? 1 : 0;
            const long rho = (c ? u - a.tiles[0] : u) * 128 + row;
            const int n_c = a.n_cls[c];
            const bool valid = rho < (long)a.n_walkers * n_c;
            long b = 0; int i = 0, j = 0;
            float dist = 0.f;
            if (valid) {
                b = rho / n_c;
                const int pk = pair_tab[(c ? a.n_cls[0] : 0) + (int)(rho - b * n_c)];
                i = pk >> 8; j = pk & 255;
                const float *ri = a.r + (b * N + i) * 3, *rj = a.r + (b * N + j) * 3;
                const float dx = rj[0] - ri[0], dy = rj[1] - ri[1], dz = rj[2] - ri[2];
                dist = i == j ? 0.f : sqrtf(dx * dx + dy * dy + dz * dz);
            }
            const long o_ij = ((b * N + i) * N + j) * 32L, o_ji = ((b * N + j) * N + i) * 32L;
            auto store_w = [&](int it, const float (&w)[32]) {
                if (!valid) return;
                float4 *p = reinterpret_cast<float4 *>(a.out[it] + o_ij);
#pragma unroll
                for (int q = 0; q < 8; ++q) p[q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                if (i != j) {
                    float4 *p2 = reinterpret_cast<float4 *>(a.out[it] + o_ji);
#pragma unroll
                    for (int q = 0; q < 8; ++q) p2[q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                }
            };
            float x[32];
            {   // iteration 0: one input feature
                float w[32];
#pragma unroll
                for (int f = 0; f < 32; ++f) w[f] = tanhf(dist * w0[c * 32 + f] + bias_w[c * 32 + f]);
                store_w(0, w);
#pragma unroll
                for (int f = 0; f < 32; ++f) x[f] = tanhf(dist * h0[c * 32 + f] + bias_h[c * 32 + f]);
            }
            for (int it = 1; it < a.n_iter; ++it) {
                // this row of the operand tile, tf32 hi / lo
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 h, l;
                        const float *xv = x + kb * 16 + q * 4;
                        h.x = rna_tf32(xv[0]); h.y = rna_tf32(xv[1]); h.z = rna_tf32(xv[2]); h.w = rna_tf32(xv[3]);
                        l.x = rna_tf32(xv[0] - h.x); l.y = rna_tf32(xv[1] - h.y); l.z = rna_tf32(xv[2] - h.z); l.w = rna_tf32(xv[3] - h.w);
                        const uint32_t off = kb * (128 * TC_ROWB) + sw64_offset(row, q * 4);
                        *reinterpret_cast<float4 *>(ah + off) = h;
                        *reinterpret_cast<float4 *>(al + off) = l;
                    }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();                      // the tcgen05.ld of the previous layer are complete (wait::ld) before the accumulator is reused
                mbar_arrive(&bar_a[g]);
                mbar_wait(&bar_d[g], ph);
                ph ^= 1;
                tc_fence_after();
                const bool has_h = it + 1 < a.n_iter;
                uint32_t v0[16], v1[16];
                tmem_ld16(taddr, v0); tmem_ld16(taddr + 16, v1);
                tmem_ld_wait();
                float w[32];
                const float *bw = bias_w + (it * 2 + c) * 32;
#pragma unroll
                for (int f = 0; f < 16; ++f) {
                    w[f] = tanhf(__fmul_rn(__uint_as_float(v0[f]), a.corr) + bw[f]);
                    w[16 + f] = tanhf(__fmul_rn(__uint_as_float(v1[f]), a.corr) + bw[16 + f]);
                }
                store_w(it, w);
                if (has_h) {                             // warp-uniform
                    tmem_ld16(taddr + 32, v0); tmem_ld16(taddr + 48, v1);
                    tmem_ld_wait();
                    const float *bh = bias_h + (it * 2 + c) * 32;
#pragma unroll
                    for (int f = 0; f < 16; ++f) {
                        x[f] = (x[f] + tanhf(__fmul_rn(__uint_as_float(v0[f]), a.corr) + bh[f])) * 0.70710678118654752f;
                        x[16 + f] = (x[16 + f] + tanhf(__fmul_rn(__uint_as_float(v1[f]), a.corr) + bh[16 + f])) * 0.70710678118654752f;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// Returns DPE_ERR_UNSUPPORTED when the model's pair layers do not have the 1 -> 32 -> 32 ... shape the kernel is written for.
int launch_pair_stream_tc(dpe_model *m, const float *r, int Bc, float *pw_base, const size_t *pw_off, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    static const bool off = getenv("DPE_PAIR_TC") && getenv("DPE_PAIR_TC")[0] == '0';
    if (off || m->gemm_path != 1 || d.emb_dim != 32 || d.n_iterations < 2 || d.n_el > 255) return DPE_ERR_UNSUPPORTED;
    for (int it = 0; it < d.n_iterations; ++it)
        if (m->it[it].dP != (it == 0 ? 1 : 32)) return DPE_ERR_UNSUPPORTED;
    PairTcArgs a;
    a.r = r; a.n_iter = d.n_iterations; a.N = d.n_el; a.U = d.n_up; a.n_walkers = Bc;
    for (int it = 0; it < d.n_iterations; ++it) {
        const IterParams &p = m->it[it];
        a.out[it] = pw_base + pw_off[it];
        a.ww[it][0] = p.w_same.w; a.wb[it][0] = p.w_same.b; a.ww[it][1] = p.w_diff.w; a.wb[it][1] = p.w_diff.b;
        a.hw[it][0] = p.h_same.w; a.hb[it][0] = p.h_same.b; a.hw[it][1] = p.h_diff.w; a.hb[it][1] = p.h_diff.b;
    }
    const int U = d.n_up, D = d.n_el - U;
    a.n_cls[0] = U * (U + 1) / 2 + D * (D + 1) / 2;
    a.n_cls[1] = U * D;
    for (int c = 0; c < 2; ++c) a.tiles[c] = ((long)Bc * a.n_cls[c] + 127) / 128;
    a.corr = tc_rz_comp(32);
    const int n_l = d.n_iterations - 1;
    const size_t smem = 1024 + (size_t)n_l * 4 * PT_W_BYTES + (size_t)PT_GROUPS * 2 * PT_A_BYTES + (size_t)(d.n_iterations * 128 + 128) * sizeof(float) +
                        (size_t)((a.n_cls[0] + a.n_cls[1] + 1) & ~1) * sizeof(int) + 2 * PT_GROUPS * sizeof(uint64_t) + 64;
    if (smem > (size_t)DPE_SMEM_OPTIN) return DPE_ERR_UNSUPPORTED;
    if (int e = opt_in_smem(m, KID_PAIR_TC, k_pair_stream_tc)) return e;
    const long n_tiles = a.tiles[0] + a.tiles[1];
    long grid = (n_tiles + PT_GROUPS - 1) / PT_GROUPS;
    if (grid > m->n_sm) grid = m->n_sm;
    k_pair_stream_tc<<<(int)grid, PT_THREADS, smem, s>>>(a);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

}  // namespace dpe
