// Dense layers on the 5th-generation tensor cores: FP32-accurate 3xTF32 GEMM with tcgen05.mma, TMEM
// accumulators and TMA-staged operands (sm_100a).
//
//   C[m, n] = sum_k X[m, k] W[k, n]      X: activations (rows = walker x electron x channel), W: layer weights
//
// The product is computed transposed, D^T[n, m] = sum_k Wt[n, k] X[m, k]:
//   * MMA "A" operand (M side, 128 lanes)  = Wt  [N_out, K]  K-major  -> TMEM lane  = output feature
//   * MMA "B" operand (N side, <=256 cols) = X   [rows,  K]  K-major  -> TMEM column = activation row
// so that every epilogue thread owns one output feature and walks over rows: a warp stores 32 consecutive
// features of one row (128 B, coalesced), and a later fusion of the tanh/tangent/Laplacian rule needs no
// cross-thread reduction (the channels of one electron are consecutive columns).
//
// FP32 accuracy (the reference forces true FP32, process_molecule.py:20-29): x = xh + xl, w = wh + wl with
// xh = rna_tf32(x), xl = rna_tf32(x - xh);  x*w ~= wh*xh + wh*xl + wl*xh   (3 tcgen05.mma.kind::tf32, FP32 accumulate).
// W is split once per parameter update (k_split_transpose); X is split in shared memory by a splitter warpgroup
// between the TMA arrival and the MMA issue (an element-wise rewrite, so it is oblivious to the 64B swizzle).
//
// CTA = 14 warps, persistent (one per SM), tile = 256 features x NMMA (<=256) rows, K in slabs of 16 floats:
//   warp 0      TMA producer (Wt_hi, Wt_lo, X slabs -> 3-stage ring, mbarrier complete_tx)
//   warp 1      TMEM allocation + MMA issuer (12 MMAs per slab: 2 k-steps x 2 feature halves x 3 products)
//   warps 2-5   splitter (X -> xh in place, xl into its own buffer, fence.proxy.async)
//   warps 6-13  epilogue (tcgen05.ld 32x32b -> coalesced global stores)
// TMEM: two accumulators [128 lanes x 256 columns] (features 0-127 / 128-255) = all 512 columns.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>
#include "dpe_internal.cuh"
#include "tc_common.cuh"

namespace dpe {

constexpr int TC_STAGES = 3;
constexpr int TC_FEAT = 256;              // features per tile (2 x M=128)
constexpr int TC_W_BYTES = TC_FEAT * TC_ROWB;          // 16 KB per hi / lo
constexpr int TC_X_BYTES = 256 * TC_ROWB;              // 16 KB (sized for NMMA = 256)
constexpr int TC_STAGE_BYTES = 2 * TC_W_BYTES + 2 * TC_X_BYTES;   // 64 KB
constexpr int TC_OUT_ROWS = 16;                                   // rows per TMA-store box (= one tcgen05.ld x16 chunk)
constexpr int TC_OUT_BYTES = TC_OUT_ROWS * 128 * 4;               // 8 KB: 16 rows x 128 features
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*barriers*/ + 4 * TC_OUT_BYTES /*store staging*/ + 1024 /*alignment slack*/;
constexpr int TC_THREADS = 14 * 32;

struct TcArgs {
    float *C;
    int ldc, c_seg_stride, c_seg_off, c_col_off;   // row of segment s, local row m -> s * c_seg_stride + c_seg_off + m
    int n_seg, seg_len;                            // rows per segment
    int N_out, K;
    int nmma;                                      // MMA N = TMA box rows (multiple of 16, <= 256)
    int tile_rows;                                 // rows a tile owns (<= nmma; a multiple of the channel count when epi != 0)
    int n_rt, n_ft;                                // row tiles per segment, feature tiles
    // fused epilogue. 0: plain store. 1: bias + spin-mean addend + tanh rule of the layer (what k_act does in a second pass).
    // 2: envelope multiply of the backflow GEMM (envelope_orbitals.py:96-127) with the product rule
    int epi, nch;                                  // channels per (walker, electron) group
    const float *r, *R, *spa, *envw;               // walker positions [n_seg, n_el, 3], ions [I,3], softplus(alpha) / weights [I, N_out]
    int n_el, n_ion, el_base;                      // electron index of local group 0 of a segment (0 for spin-up, n_up for spin-down)
    const float *bias, *add; int gpa;              // epi == 1: bias [N_out], addend [(group / gpa)][nch][N_out]
    int pipe;                                      // software-pipelined TMEM loads in the epilogue
    int spt;                                       // segments per tile (> 1: short segments, e.g. the spin blocks of a forward pass, are packed into one tile)
    int tma_store;                                 // plain epilogue through shared-memory staging + TMA tensor stores
    int n_st;                                      // pipeline stages of the CTA-pair kernel
    float corr;                                    // accumulation-bias compensation factor, see tc_rz_compensation()
    int seg_split;                                 // fused epilogues: channel segments per (walker, electron) group (work items per group)
    int add_smem;                                  // epi 1: the addend rows of a tile are TMA-prefetched into shared memory (else read from global)
    int w_seg_k;                                   // K offset of the W maps per segment (split-K products: a segment is a K chunk), else 0
    long long *tl;                                 // debug timeline (DPE_GEMM_TIMELINE): [tile][4] clock64 stamps of CTA 0, or nullptr
};

constexpr int TC_TL_TILES = 64;
#define TC_TL(role, idx) do { if (a.tl && blockIdx.x == 0 && (idx) < TC_TL_TILES) a.tl[(idx) * 8 + (role)] = clock64(); } while (0)

// ---------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc_3xtf32(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                 const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_c, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t *bar_full = bars;                    // [S] TMA landed
    uint64_t *bar_split = bars + TC_STAGES;       // [S] splitter done
    uint64_t *bar_empty = bars + 2 * TC_STAGES;   // [S] MMAs that read the stage retired
    uint64_t *bar_tfull = bars + 3 * TC_STAGES;   // accumulators complete
    uint64_t *bar_tempty = bar_tfull + 1;         // accumulators drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_tempty + 1);
    uint8_t *out_stage = smem + TC_STAGES * TC_STAGE_BYTES + 1024;     // [feature half][2][16 rows][128 features] store staging

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (a.K + TC_BK - 1) / TC_BK;
    const long n_tiles = (long)((a.n_seg + a.spt - 1) / a.spt) * a.n_rt * a.n_ft;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_split[s], 128);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx = 2 * TC_W_BYTES + (a.spt > 1 ? a.tile_rows : a.nmma) * TC_ROWB;
            for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int ft = (int)(t % a.n_ft);
                const long rest = t / a.n_ft;
                const int rt = (int)(rest % a.n_rt), seg = (int)(rest / a.n_rt);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t *st = smem + stage * TC_STAGE_BYTES;
                    mbar_expect_tx(&bar_full[stage], tx);
                    tma_load_2d(st, &map_wh, &bar_full[stage], kb * TC_BK + seg * a.w_seg_k, ft * TC_FEAT);
                    tma_load_2d(st + TC_W_BYTES, &map_wl, &bar_full[stage], kb * TC_BK + seg * a.w_seg_k, ft * TC_FEAT);
                    tma_load_3d(st + 2 * TC_W_BYTES, &map_x, &bar_full[stage], kb * TC_BK, a.spt > 1 ? 0 : rt * a.tile_rows, seg * a.spt);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.nmma >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0, tcount = 0; uint32_t phase = 0, tphase = 0;
            for (long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
                mbar_wait(bar_tempty, tphase ^ 1);
                tc_fence_after();
                TC_TL(0, tcount);                                   // accumulators free: first MMA of the tile can issue
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_split[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * TC_STAGE_BYTES);
                    const uint32_t wh = st, wl = st + TC_W_BYTES, xh = st + 2 * TC_W_BYTES, xl = st + 2 * TC_W_BYTES + TC_X_BYTES;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        const uint32_t ko = kk * 32;               // 8 tf32 = 32 bytes along K inside the swizzle atom
                        const uint64_t dxh = make_desc_sw64(xh + ko), dxl = make_desc_sw64(xl + ko);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t d = tmem_base + h * 256;
                            const uint64_t dwh = make_desc_sw64(wh + h * 128 * TC_ROWB + ko);
                            const uint64_t dwl = make_desc_sw64(wl + h * 128 * TC_ROWB + ko);
                            tc_mma_tf32(d, dwl, dxh, idesc, (kb | kk) ? 1u : 0u);   // small terms first
                            tc_mma_tf32(d, dwh, dxl, idesc, 1u);
                            tc_mma_tf32(d, dwh, dxh, idesc, 1u);
                        }
                    }
                    tc_commit(&bar_empty[stage]);                   // frees the stage once these MMAs retire
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(bar_tfull);
                TC_TL(1, tcount);                                   // last MMA of the tile issued
                tphase ^= 1;
            }
        }
    } else if (warp < 6) {
        // ================================ splitter ================================
        const int tid = threadIdx.x - 64;       // 0..127
        int stage = 0; uint32_t phase = 0;
        const int n_f4 = a.nmma * (TC_ROWB / 16);
        for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                float4 *xh = reinterpret_cast<float4 *>(smem + stage * TC_STAGE_BYTES + 2 * TC_W_BYTES);
                float4 *xl = reinterpret_cast<float4 *>(smem + stage * TC_STAGE_BYTES + 2 * TC_W_BYTES + TC_X_BYTES);
                for (int base = tid; base < n_f4; base += 4 * 128) {      // loads batched ahead of the stores (they may alias for the compiler)
                    float4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (base + k * 128 < n_f4) v[k] = xh[base + k * 128];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (base + k * 128 < n_f4) {
                            float4 h, l;
                            h.x = rna_tf32(v[k].x); h.y = rna_tf32(v[k].y); h.z = rna_tf32(v[k].z); h.w = rna_tf32(v[k].w);
                            l.x = rna_tf32(v[k].x - h.x); l.y = rna_tf32(v[k].y - h.y); l.z = rna_tf32(v[k].z - h.z); l.w = rna_tf32(v[k].w - h.w);
                            xh[base + k * 128] = h;
                            xl[base + k * 128] = l;
                        }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&bar_split[stage]);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue ================================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int h = (warp - 6) >> 2;           // accumulator (feature half)
        uint32_t tphase = 0;
        int tcount = 0, chunk = 0;
        for (long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tcount) {
            const int ft = (int)(t % a.n_ft);
            const long rest = t / a.n_ft;
            const int rt = (int)(rest % a.n_rt), seg = (int)(rest / a.n_rt);
            const int f = ft * TC_FEAT + h * 128 + q * 32 + lane;
            const int m0 = rt * a.tile_rows;
            const int rows_valid = min(a.tile_rows, a.seg_len - m0);
            float *cbase = a.C + ((long)seg * a.c_seg_stride + a.c_seg_off + m0) * a.ldc + a.c_col_off + f;
            const bool f_ok = f < a.N_out;
            // envelope state (epi == 2): group = (walker seg, electron el_base + (m0 + col) / C), channel cch
            int cch = 0, grp = a.epi == 2 ? m0 / a.nch : 0;
            float env = 1.f, e1x = 0.f, e1y = 0.f, e1z = 0.f, el = 0.f, bf0 = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
            int ci = 0;
            // activation state (epi == 1): tanh value, 1 - tanh^2, running sum of squared tangents of the current group
            float act_y = 0.f, act_d1 = 0.f, act_ssq = 0.f;
            const float act_b = (a.epi == 1 && f_ok && a.bias) ? a.bias[f] : 0.f;
            long agrp = a.epi == 1 ? ((long)seg * a.seg_len + m0) / a.nch : 0;     // global (walker, electron) group of column 0
            mbar_wait(bar_tfull, tphase);
            tc_fence_after();
            if (threadIdx.x == 6 * 32) TC_TL(2, tcount);           // MMAs retired: epilogue starts
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * 256;
            // one 16-column chunk of this thread's feature: plain store or fused envelope (epi == 2)
            auto process = [&](const uint32_t (&v)[16], int c0) {
                if (a.epi == 0 && a.spt > 1) {
                    if (f_ok) {       // packed short segments: column -> (segment, row in segment)
                        int sl = c0 / a.seg_len, mm = c0 - sl * a.seg_len;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (c0 + j < a.tile_rows && seg * a.spt + sl < a.n_seg)
                                a.C[((long)(seg * a.spt + sl) * a.c_seg_stride + a.c_seg_off + mm) * a.ldc + a.c_col_off + f] = __fmul_rn(__uint_as_float(v[j]), a.corr);
                            if (++mm == a.seg_len) { mm = 0; ++sl; }
                        }
                    }
                } else if (a.epi == 0) {
                    if (f_ok) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < rows_valid) cbase[(long)(c0 + j) * a.ldc] = __fmul_rn(__uint_as_float(v[j]), a.corr);
                    }
                } else if (a.epi == 1) {
                    // z = x W + bias + addend;  y = tanh(z_0), t_k = (1 - y^2) z_k, lap = (1 - y^2) z_lap - 2 y (1 - y^2) sum_k z_k^2
                    // (same arithmetic and order as k_act).  The addends of the chunk are fetched ahead of the dependent chain.
                    float ad[16];
                    {
                        int cc = cch; long gg = agrp;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            ad[j] = (a.add && f_ok && c0 + j < rows_valid) ? a.add[((gg / a.gpa) * a.nch + cc) * (long)a.N_out + f] : 0.f;
                            if (++cc == a.nch) { cc = 0; ++gg; }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (c0 + j < rows_valid) {
                            float z = __fmul_rn(__uint_as_float(v[j]), a.corr);
                            float o;
                            if (cch == 0) {
                                z += act_b;
                                z += ad[j];
                                act_y = tanhf(z);
                                act_d1 = 1.f - act_y * act_y;
                                act_ssq = 0.f;
                                o = act_y;
                            } else if (cch < a.nch - 1) {
                                z += ad[j];
                                act_ssq = fmaf(z, z, act_ssq);
                                o = act_d1 * z;
                            } else {
                                z += ad[j];
                                o = act_d1 * z - 2.f * act_y * act_d1 * act_ssq;
                            }
                            if (f_ok) cbase[(long)(c0 + j) * a.ldc] = o;
                            if (++cch == a.nch) { cch = 0; ++agrp; }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (c0 + j < rows_valid) {
                            const float val = __fmul_rn(__uint_as_float(v[j]), a.corr);
                            if (cch == 0) {
                                const int i = a.el_base + grp;
                                ci = 1 + 3 * i;
                                bf0 = val;
                                env = 0.f; e1x = e1y = e1z = el = 0.f;
                                if (f_ok) {
                                    const float *ri = a.r + ((long)seg * a.n_el + i) * 3;
                                    const float rx = ri[0], ry = ri[1], rz = ri[2];
                                    for (int J = 0; J < a.n_ion; ++J) {
                                        const float dx = rx - a.R[J * 3], dy = ry - a.R[J * 3 + 1], dz = rz - a.R[J * 3 + 2];
                                        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                                        const float al = a.spa[(long)J * a.N_out + f];
                                        const float e = __fmul_rn(a.envw[(long)J * a.N_out + f], expf(-al * d));
                                        env = __fadd_rn(env, e);
                                        if (a.nch > 1) {
                                            const float inv = 1.f / d, g = -al * e * inv;
                                            e1x = fmaf(g, dx, e1x); e1y = fmaf(g, dy, e1y); e1z = fmaf(g, dz, e1z);
                                            el = fmaf(e, al * al - 2.f * al * inv, el);
                                        }
                                    }
                                }
                            }
                            float o = val * env;
                            if (a.nch > 1) {
                                if (cch == ci) { tx = val; o += e1x * bf0; }
                                else if (cch == ci + 1) { ty = val; o += e1y * bf0; }
                                else if (cch == ci + 2) { tz = val; o += e1z * bf0; }
                                else if (cch == a.nch - 1) o += el * bf0 + 2.f * (e1x * tx + e1y * ty + e1z * tz);
                            }
                            if (f_ok) cbase[(long)(c0 + j) * a.ldc] = o;
                            if (++cch == a.nch) { cch = 0; ++grp; }
                        }
                    }
                }
            };
            if (a.tma_store) {
                // Plain store: the accumulator chunk [16 rows x 128 features of this half] is transposed through shared memory
                // and leaves as one TMA tensor store (bounds clipped by the tensor map).  The epilogue warps only copy
                // TMEM -> registers -> shared, so the accumulators are released ~5x earlier than with per-thread global stores
                // (timeline: 22 k of 56 k clocks per tile were exposed epilogue), and the stores overlap the next tile's MMAs.
                float *stage_base = reinterpret_cast<float *>(out_stage + h * 2 * TC_OUT_BYTES);
                const bool elected = (q == 0 && lane == 0);
                const int bar_id = 2 + h;
                uint32_t v[16];
                for (int c0 = 0; c0 < a.nmma; c0 += 16, ++chunk) {      // `chunk` runs on across tiles: staging buffers strictly alternate
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (c0 + 16 >= a.nmma) { tc_fence_before(); mbar_arrive(bar_tempty); }
                    if (elected) tma_store_wait_read<1>();              // the store that last read this staging buffer is done with it
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    float *st = stage_base + (chunk & 1) * (TC_OUT_BYTES / 4) + q * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 16; ++j) st[j * 128] = __fmul_rn(__uint_as_float(v[j]), a.corr);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (elected) {
                        if (c0 < rows_valid) tma_store_3d(&map_c, stage_base + (chunk & 1) * (TC_OUT_BYTES / 4), ft * TC_FEAT + h * 128, m0 + c0, seg);
                        tma_store_commit();                             // always: one group per chunk keeps wait_group.read 1 aligned
                    }
                }
                if (threadIdx.x == 6 * 32) TC_TL(3, tcount);
                tphase ^= 1;
                continue;
            }
            // TMEM -> registers is software pipelined: the next chunk travels while the current one is stored
            uint32_t v0[16], v1[16];
            if (!a.pipe) {
                for (int c0 = 0; c0 < a.nmma; c0 += 16) {
                    tmem_ld16(taddr + c0, v0);
                    tmem_ld_wait();
                    if (c0 + 16 >= a.nmma) { tc_fence_before(); mbar_arrive(bar_tempty); }
                    process(v0, c0);
                }
            } else {
            tmem_ld16(taddr, v0);
            for (int c0 = 0; c0 < a.nmma; c0 += 32) {
                tmem_ld_wait();
                if (c0 + 16 < a.nmma) tmem_ld16(taddr + c0 + 16, v1);
                else { tc_fence_before(); mbar_arrive(bar_tempty); }   // everything is in registers: accumulators are free
                process(v0, c0);
                if (c0 + 16 < a.nmma) {
                    tmem_ld_wait();
                    if (c0 + 32 < a.nmma) tmem_ld16(taddr + c0 + 32, v0);
                    else { tc_fence_before(); mbar_arrive(bar_tempty); }
                    process(v1, c0 + 16);
                }
            }
            }
            if (threadIdx.x == 6 * 32) TC_TL(3, tcount);           // this warp's share of the tile stored
            tphase ^= 1;
        }
        if (a.tma_store && q == 0 && lane == 0) tma_store_wait_all();     // every tensor store of this thread has landed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ----------------------------------------------------------------------------------------------------------------
// Narrow layers (N_out <= 32: the h_map layer of the SchNet convolution, ferminet_embedding.py:156-158).
// Here the roles are the natural ones, D[m, n] = sum_k X[m, k] Wt[n, k]: MMA M side = 128 activation rows (TMEM lanes),
// N side = 32 output features, so no tensor lane is wasted on padding features.  Wt_hi / Wt_lo (<= 2 x 40 KB) stay resident
// in shared memory for the whole kernel; X streams through a 3-stage ring exactly as in the wide kernel (TMA -> splitter
// -> 12 MMAs of M128 x N32 x K8 per slab); accumulators are double buffered in TMEM (2 x 64 columns), so the epilogue
// (one thread = one output row = 128 contiguous bytes) overlaps the MMAs of the next tile.
// ----------------------------------------------------------------------------------------------------------------
constexpr int TR_ROWS = 256;                       // rows per tile (2 x M=128)
constexpr int TR_N = 32;                           // MMA N
constexpr int TR_STAGES = 6;                       // upper bound; the launcher picks a.n_st so that stages + resident W fit
constexpr int TR_X_BYTES = TR_ROWS * TC_ROWB;      // 16 KB
constexpr int TR_STAGE_BYTES = 2 * TR_X_BYTES;     // raw->hi, lo
constexpr int TR_WSLAB = TR_N * TC_ROWB;           // 2 KB per K slab per hi / lo
constexpr int TR_MAX_KB = 20;                      // K <= 320
constexpr int TR_SMEM_BYTES = TR_STAGES * TR_STAGE_BYTES + 2 * TR_MAX_KB * TR_WSLAB + 1024 + 1024;
constexpr int TR_THREADS = 10 * 32;

struct TrArgs {
    float *C; int ldc, c_col_off;
    int M, N_out, K;
    float corr;        // accumulation-bias compensation factor, see tc_rz_compensation()
    int n_st;          // X stages (the kernel is HBM-latency bound: as many as fit next to the resident weights)
    const float *bias; // forward pass: y = tanh(x W + bias) in the epilogue (nullptr: plain store)
};

__global__ void __launch_bounds__(TR_THREADS, 1)
k_gemm_tc_rows_3xtf32(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                      const __grid_constant__ CUtensorMap map_wl, TrArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    const int n_st = a.n_st, n_kbw = (a.K + TC_BK - 1) / TC_BK;
    uint8_t *w_hi = smem + n_st * TR_STAGE_BYTES;               // [n_kb][32 rows x 64 B]
    uint8_t *w_lo = w_hi + n_kbw * TR_WSLAB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(w_lo + n_kbw * TR_WSLAB);
    uint64_t *bar_full = bars, *bar_split = bars + TR_STAGES, *bar_empty = bars + 2 * TR_STAGES;
    uint64_t *bar_tfull = bars + 3 * TR_STAGES;                 // [2]
    uint64_t *bar_tempty = bar_tfull + 2;                       // [2]
    uint64_t *bar_w = bar_tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_w + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (a.K + TC_BK - 1) / TC_BK;
    const long n_tiles = ((long)a.M + TR_ROWS - 1) / TR_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TR_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_split[s], 128); mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], 128); }
        mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(bar_w, 2 * n_kb * TR_WSLAB);
            for (int kb = 0; kb < n_kb; ++kb) {
                tma_load_2d(w_hi + kb * TR_WSLAB, &map_wh, bar_w, kb * TC_BK, 0);
                tma_load_2d(w_lo + kb * TR_WSLAB, &map_wl, bar_w, kb * TC_BK, 0);
            }
            int stage = 0; uint32_t phase = 0;
            for (long t = blockIdx.x; t < n_tiles; t += gridDim.x)
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    mbar_expect_tx(&bar_full[stage], TR_X_BYTES);
                    tma_load_3d(smem + stage * TR_STAGE_BYTES, &map_x, &bar_full[stage], kb * TC_BK, (int)(t * TR_ROWS), 0);
                    if (++stage == n_st) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TR_N >> 3) << 17) | ((128u >> 4) << 24);
            mbar_wait(bar_w, 0);
            int stage = 0; uint32_t phase = 0, tph0 = 0, tph1 = 0;
            int buf = 0;
            for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&bar_tempty[buf], (buf ? tph1 : tph0) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_split[stage], phase);
                    tc_fence_after();
                    const uint32_t xh = smem_u32(smem + stage * TR_STAGE_BYTES), xl = xh + TR_X_BYTES;
                    const uint32_t wh = smem_u32(w_hi + kb * TR_WSLAB), wl = smem_u32(w_lo + kb * TR_WSLAB);
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        const uint32_t ko = kk * 32;
                        const uint64_t dwh = make_desc_sw64(wh + ko), dwl = make_desc_sw64(wl + ko);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t d = tmem_base + buf * 64 + h * TR_N;
                            const uint64_t dxh = make_desc_sw64(xh + h * 128 * TC_ROWB + ko), dxl = make_desc_sw64(xl + h * 128 * TC_ROWB + ko);
                            tc_mma_tf32(d, dxh, dwl, idesc, (kb | kk) ? 1u : 0u);
                            tc_mma_tf32(d, dxl, dwh, idesc, 1u);
                            tc_mma_tf32(d, dxh, dwh, idesc, 1u);
                        }
                    }
                    tc_commit(&bar_empty[stage]);
                    if (++stage == n_st) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bar_tfull[buf]);
                if (buf) tph1 ^= 1; else tph0 ^= 1;
                buf ^= 1;
            }
        }
    } else if (warp < 6) {
        const int tid = threadIdx.x - 64;
        int stage = 0; uint32_t phase = 0;
        for (long t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                float4 *xh = reinterpret_cast<float4 *>(smem + stage * TR_STAGE_BYTES);
                float4 *xl = reinterpret_cast<float4 *>(smem + stage * TR_STAGE_BYTES + TR_X_BYTES);
#pragma unroll
                for (int i = tid; i < TR_X_BYTES / 16; i += 128) {
                    float4 v = xh[i], h, l;
                    h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
                    l.x = rna_tf32(v.x - h.x); l.y = rna_tf32(v.y - h.y); l.z = rna_tf32(v.z - h.z); l.w = rna_tf32(v.w - h.w);
                    xh[i] = h;
                    xl[i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&bar_split[stage]);
                if (++stage == n_st) { stage = 0; phase ^= 1; }
            }
    } else {
        const int q = warp & 3;
        uint32_t tph0 = 0, tph1 = 0;
        int buf = 0;
        for (long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(&bar_tfull[buf], buf ? tph1 : tph0);
            tc_fence_after();
            uint32_t v0[16], v1[16], v2[16], v3[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64;
            tmem_ld16(taddr, v0); tmem_ld16(taddr + 16, v1); tmem_ld16(taddr + 32, v2); tmem_ld16(taddr + 48, v3);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bar_tempty[buf]);
            auto store_row = [&](long row, const uint32_t (&lo)[16], const uint32_t (&hi)[16]) {
                if (row >= a.M) return;
                float *dst = a.C + row * a.ldc + a.c_col_off;
                float o[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    o[j] = __fmul_rn(__uint_as_float(lo[j]), a.corr);
                    o[16 + j] = __fmul_rn(__uint_as_float(hi[j]), a.corr);
                }
                if (a.bias) {                              // forward pass: bias + tanh here instead of a separate pass over the output
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = j < a.N_out ? tanhf(o[j] + a.bias[j]) : 0.f;
                }
                if (a.N_out == 32 && ((a.ldc | a.c_col_off) & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) *reinterpret_cast<float4 *>(dst + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < a.N_out) dst[j] = o[j];
                }
            };
            const long r0 = t * TR_ROWS + q * 32 + lane;
            store_row(r0, v0, v1);            // rows 0..127 of the tile (accumulator columns 0..31)
            store_row(r0 + 128, v2, v3);      // rows 128..255 (columns 32..63)
            if (buf) tph1 ^= 1; else tph0 ^= 1;
            buf ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}


// ----------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the wide kernel (tcgen05 cta_group::2): two SMs of a TPC share one 256-feature x nmma-row tile.
// CTA r of the pair holds features [128 r, 128 r + 128) of Wt_hi / Wt_lo and rows [r nmma/2, (r+1) nmma/2) of X in its own
// shared memory; the leader (rank 0) issues M256 x N(nmma) x K8 MMAs that read both halves, and every SM accumulates ITS 128
// features x nmma rows in its own TMEM.  That is 256 TMEM columns per tile instead of 512, so the accumulators are double
// buffered and the epilogue (TMEM -> staging -> TMA stores) runs under the next tile's MMAs; per SM the X operand traffic
// (TMA writes, splitter passes, MMA operand reads) is halved.
//   warp 0      TMA producer (own halves)            warp 1      TMEM allocation; in the leader: MMA issuer
//   warps 2-5   splitter (own X half)                warps 6-13  epilogue (own accumulator)
// Cross-CTA signalling: splitters and epilogues of both CTAs arrive on the LEADER's mbarriers (mapa + .shared::cluster);
// tcgen05.commit multicasts "stage free" / "accumulator full" to both CTAs.
// ----------------------------------------------------------------------------------------------------------------
constexpr int T2_STAGES = 6;                                     // upper bound; the launcher picks a.n_st <= T2_STAGES that fits
constexpr int T2_W_BYTES = 128 * TC_ROWB;                         // 8 KB per hi / lo (128 features)
constexpr int T2_X_BYTES = 128 * TC_ROWB;                         // 8 KB (128 rows)
constexpr int T2_STAGE_BYTES = 2 * T2_W_BYTES + 2 * T2_X_BYTES;   // 32 KB
constexpr int T2_PLAIN_STAGES = 6;
constexpr int T2_SMEM_BYTES = T2_PLAIN_STAGES * T2_STAGE_BYTES + 1024 + 4 * TC_OUT_BYTES + 1024;

constexpr int T2_FUSED_EPI_WARPS = 16;             // fused epilogue: four warps per TMEM lane quarter, one (walker, electron) group each
template <bool FUSED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FUSED ? (6 + T2_FUSED_EPI_WARPS) * 32 : TC_THREADS, 1)
k_gemm_tc2_3xtf32(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                  const __grid_constant__ CUtensorMap map_wl, const __grid_constant__ CUtensorMap map_c,
                  const __grid_constant__ CUtensorMap map_add, TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n_st = a.n_st;                           // the fused epilogue may trade a stage for its addend buffers
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + n_st * T2_STAGE_BYTES);
    uint64_t *bar_full = bars;                      // [S] own TMA landed
    uint64_t *bar_split = bars + T2_STAGES;         // [S] leader only: both splitter groups done (one elected arrival per CTA:
                                                    //     128 remote arrivals per stage serialise on the DSMEM path)
    uint64_t *bar_empty = bars + 2 * T2_STAGES;     // [S] MMAs that read the stage retired (multicast commit)
    uint64_t *bar_tfull = bars + 3 * T2_STAGES;     // [2] accumulator complete (multicast commit)
    uint64_t *bar_tempty = bar_tfull + 2;           // [2] leader only: the four epilogue groups of the pair drained the buffer
    uint64_t *bar_afull = bar_tempty + 2;           // [2] addend tile of the fused epilogue landed (TMA)
    uint64_t *bar_aempty = bar_afull + 2;           // [2] ... and has been consumed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_aempty + 2);
    uint8_t *out_stage = smem + n_st * T2_STAGE_BYTES + 1024;                      // plain epilogue: TMA-store staging
    float *addbuf = reinterpret_cast<float *>(smem + n_st * T2_STAGE_BYTES + 1024); // fused epilogue: [2 tiles][2 walkers][nch][128 features] addend rows (add_smem), then the partial sums of the channel segments

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_kb = (a.K + TC_BK - 1) / TC_BK;
    const long n_tiles = (long)a.n_seg * a.n_rt * a.n_ft;
    const int half_rows = a.nmma >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_split[s], 2); mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], FUSED ? 2 * T2_FUSED_EPI_WARPS : 4);
            mbar_init(&bar_afull[b], 1); mbar_init(&bar_aempty[b], T2_FUSED_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                             // barriers of both CTAs initialised, TMEM allocated on both SMs
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t tx = 2 * T2_W_BYTES + half_rows * TC_ROWB;
            int tcount = 0;
            for (long t = pair; t < n_tiles; t += n_pairs, ++tcount) {
                const int ft = (int)(t % a.n_ft);
                const long rest = t / a.n_ft;
                const int rt = (int)(rest % a.n_rt), seg = (int)(rest / a.n_rt);
                if (FUSED && a.add && a.add_smem) {
                    // addends of this tile: one box of nch rows x 128 features per walker the tile touches (at most two)
                    const int slot = tcount & 1, m0 = rt * a.tile_rows;
                    const int rows_valid = min(a.tile_rows, a.seg_len - m0);
                    const long g_first = ((long)seg * a.seg_len + m0) / a.nch, g_last = g_first + (rows_valid + a.nch - 1) / a.nch - 1;
                    const long w_first = g_first / a.gpa;
                    const int n_w = (int)(g_last / a.gpa - w_first) + 1;
                    mbar_wait(&bar_aempty[slot], (((uint32_t)tcount >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(&bar_afull[slot], (uint32_t)(n_w * a.nch * 512));
                    for (int w = 0; w < n_w; ++w)
                        tma_load_2d(addbuf + (size_t)(slot * 2 + w) * a.nch * 128, &map_add, &bar_afull[slot], ft * TC_FEAT + (int)rank * 128,
                                    (int)((w_first + w) * a.nch));
                }
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t *st = smem + stage * T2_STAGE_BYTES;
                    mbar_expect_tx(&bar_full[stage], tx);
                    tma_load_2d(st, &map_wh, &bar_full[stage], kb * TC_BK + seg * a.w_seg_k, ft * TC_FEAT + (int)rank * 128);
                    tma_load_2d(st + T2_W_BYTES, &map_wl, &bar_full[stage], kb * TC_BK + seg * a.w_seg_k, ft * TC_FEAT + (int)rank * 128);
                    tma_load_3d(st + 2 * T2_W_BYTES, &map_x, &bar_full[stage], kb * TC_BK, rt * a.tile_rows + (int)rank * half_rows, seg);
                    if (++stage == n_st) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.nmma >> 3) << 17) | ((256u >> 4) << 24);
            int stage = 0; uint32_t phase = 0, tph0 = 0, tph1 = 0;
            int buf = 0, tcount = 0;
            for (long t = pair; t < n_tiles; t += n_pairs, ++tcount) {
                mbar_wait_cluster(&bar_tempty[buf], (buf ? tph1 : tph0) ^ 1);
                tc_fence_after();
                TC_TL(0, tcount);
                const uint32_t d = tmem_base + buf * 256;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait_cluster(&bar_split[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * T2_STAGE_BYTES);
                    const uint32_t wh = st, wl = st + T2_W_BYTES, xh = st + 2 * T2_W_BYTES, xl = xh + T2_X_BYTES;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        const uint32_t ko = kk * 32;
                        const uint64_t dxh = make_desc_sw64(xh + ko), dxl = make_desc_sw64(xl + ko);
                        const uint64_t dwh = make_desc_sw64(wh + ko), dwl = make_desc_sw64(wl + ko);
                        tc_mma2_tf32(d, dwl, dxh, idesc, (kb | kk) ? 1u : 0u);   // small terms first
                        tc_mma2_tf32(d, dwh, dxl, idesc, 1u);
                        tc_mma2_tf32(d, dwh, dxh, idesc, 1u);
                    }
                    tc_commit2(&bar_empty[stage], 3);          // frees the stage in both CTAs once these MMAs retire
                    if (++stage == n_st) { stage = 0; phase ^= 1; }
                }
                tc_commit2(&bar_tfull[buf], 3);
                TC_TL(1, tcount);
                if (buf) tph1 ^= 1; else tph0 ^= 1;
                buf ^= 1;
            }
        }
    } else if (warp < 6) {
        const int tid = threadIdx.x - 64;
        int stage = 0; uint32_t phase = 0;
        const int n_f4 = half_rows * (TC_ROWB / 16);
        int tcount = 0;
        for (long t = pair; t < n_tiles; t += n_pairs, ++tcount) {
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                float4 *xh = reinterpret_cast<float4 *>(smem + stage * T2_STAGE_BYTES + 2 * T2_W_BYTES);
                float4 *xl = reinterpret_cast<float4 *>(smem + stage * T2_STAGE_BYTES + 2 * T2_W_BYTES + T2_X_BYTES);
                for (int base = tid; base < n_f4; base += 4 * 128) {      // loads batched ahead of the stores (they may alias for the compiler)
                    float4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (base + k * 128 < n_f4) v[k] = xh[base + k * 128];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (base + k * 128 < n_f4) {
                            float4 h, l;
                            h.x = rna_tf32(v[k].x); h.y = rna_tf32(v[k].y); h.z = rna_tf32(v[k].z); h.w = rna_tf32(v[k].w);
                            l.x = rna_tf32(v[k].x - h.x); l.y = rna_tf32(v[k].y - h.y); l.z = rna_tf32(v[k].z - h.z); l.w = rna_tf32(v[k].w - h.w);
                            xh[base + k * 128] = h;
                            xl[base + k * 128] = l;
                        }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 4, 128;" ::: "memory");
                if (tid == 0) mbar_arrive_cluster_relaxed(mapa_u32(&bar_split[stage], 0));     // the leader's barrier collects both CTAs
                if (++stage == n_st) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;                  // TMEM lane quarter = 32 features of this CTA's 128
        const int g = (warp - 6) >> 2;           // the two epilogue groups take alternate 16-row chunks
        float *stage_base = reinterpret_cast<float *>(out_stage + g * 2 * TC_OUT_BYTES);
        const bool elected = (q == 0 && lane == 0);
        const int bar_id = 2 + g;
        uint32_t tph0 = 0, tph1 = 0;
        int buf = 0, chunk = 0, tcount = 0;
        if constexpr (FUSED) {
            // Fused dense-layer epilogues (what k_act / k_envelope do in separate HBM passes); with the accumulators double buffered they
            // run under the next tile's MMAs.
            //   epi 1 (mlp.py:45-69 under the forward Laplacian):  z = x W + bias + addend;  y = tanh(z_0), t_k = (1 - y^2) z_k,
            //          lap = (1 - y^2) z_lap - 2 y (1 - y^2) sum_k z_k^2
            //   epi 2 (envelope_orbitals.py:96-127): mo = env (x) bf with the product rule
            // One thread = one output feature.  The rows of a tile are whole (walker, electron) groups of nch channels; a WORK ITEM is one
            // of a.seg_split channel segments of one group, and the four warps of a TMEM lane quarter take the items of a tile round
            // robin (a walk over rows is a dependent chain of ~180 clocks per row for a lone warp).  Splitting groups into segments keeps
            // all four warps busy when a tile holds few groups (benzene: 2 groups of 128 channels; N2: 5 groups of 44).  The only
            // state that crosses segments is sum_k z_k^2 of epi 1: the segments park their partial sums in shared memory and, after a
            // barrier of the quarter's four warps, the Laplacian rows are finished.  The spin-mean addend rows are read straight from
            // global memory (L2-resident: one walker's block serves all its electrons), fetched a chunk ahead of the dependent chain.
            const int k4 = (warp - 6) >> 2;                            // which of the four warps of this lane quarter
            const bool has_add = a.add != nullptr;
            const int S = a.seg_split, seg_ch = (a.nch + S - 1) / S;
            const bool add_smem = has_add && a.add_smem;
            float *ssq_s = addbuf + (a.add_smem ? (size_t)4 * a.nch * 128 : 0);      // [group][segment][128 features] partial sums (S > 1 only)
            const long astride = add_smem ? 128 : a.N_out;
            for (long t = pair; t < n_tiles; t += n_pairs, ++tcount) {
                const int ft = (int)(t % a.n_ft);
                const long rest = t / a.n_ft;
                const int rt = (int)(rest % a.n_rt), seg = (int)(rest / a.n_rt);
                const int m0 = rt * a.tile_rows;
                const int rows_valid = min(a.tile_rows, a.seg_len - m0);
                const int n_grp = (rows_valid + a.nch - 1) / a.nch;
                const int f = ft * TC_FEAT + (int)rank * 128 + q * 32 + lane;
                const bool f_ok = f < a.N_out;
                const int fc = f_ok ? f : 0;                               // clamped feature for loads that every lane issues
                const float act_b = (f_ok && a.bias) ? a.bias[f] : 0.f;
                float *cbase = a.C + ((long)seg * a.c_seg_stride + a.c_seg_off + m0) * a.ldc + a.c_col_off + f;
                const long agrp = ((long)seg * a.seg_len + m0) / a.nch;    // global (walker, electron) group of column 0
                const int e0 = (int)(agrp % a.gpa);                        // its electron index within the walker
                const int slot = tcount & 1;
                if (add_smem) mbar_wait(&bar_afull[slot], ((uint32_t)tcount >> 1) & 1u);
                mbar_wait(&bar_tfull[buf], buf ? tph1 : tph0);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
                const int n_items = n_grp * S;
                auto addend_rows = [&](int gi) -> const float * {
                    if (!has_add) return nullptr;
                    if (add_smem) return addbuf + ((size_t)(slot * 2 + (e0 + gi) / a.gpa) * a.nch) * 128 + q * 32 + lane;
                    return a.add + (((agrp + gi) / a.gpa) * a.nch) * (long)a.N_out + fc;
                };
                for (int item = k4; item < n_items; item += 4) {
                    const int gi = item / S, sg = item - gi * S;
                    const int col0 = gi * a.nch;
                    const int n_valid = min(a.nch, rows_valid - col0);         // rows of this group inside the segment (all of them, normally)
                    const int c_lo = sg * seg_ch, c_hi = min(n_valid, c_lo + seg_ch);
                    if (a.epi == 2) {
                        // env = sum_J w_J exp(-alpha_J |r_i - R_J|) for this (electron, orbital column); only the electron's own three tangent
                        // channels and the Laplacian channel get extra terms
                        const int i = a.el_base + m0 / a.nch + gi;
                        const int ci = 1 + 3 * i;
                        float env = 0.f, e1x = 0.f, e1y = 0.f, e1z = 0.f, el = 0.f;
                        {
                            const float *ri = a.r + ((long)seg * a.n_el + i) * 3;
                            const float rx = ri[0], ry = ri[1], rz = ri[2];
                            for (int J = 0; J < a.n_ion; ++J) {
                                const float dx = rx - a.R[J * 3], dy = ry - a.R[J * 3 + 1], dz = rz - a.R[J * 3 + 2];
                                const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                                const float al = a.spa[(long)J * a.N_out + fc];
                                const float e = __fmul_rn(a.envw[(long)J * a.N_out + fc], expf(-al * d));
                                env = __fadd_rn(env, e);
                                if (a.nch > 1) {
                                    const float inv = 1.f / d, ge = -al * e * inv;
                                    e1x = fmaf(ge, dx, e1x); e1y = fmaf(ge, dy, e1y); e1z = fmaf(ge, dz, e1z);
                                    el = fmaf(e, al * al - 2.f * al * inv, el);
                                }
                            }
                        }
                        // every row is val * env; the electron's own three tangent rows and the Laplacian row get their product-rule terms in
                        // a second, four-row patch, which keeps the walk over the rows free of per-row case distinctions
                        float *crow = cbase + (long)(col0 + c_lo) * a.ldc;
                        int cc = c_lo;
                        for (; cc + 16 <= c_hi; cc += 16) {
                            uint32_t v[16];
                            tmem_ld16(taddr + col0 + cc, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (f_ok) *crow = __fmul_rn(__uint_as_float(v[j]), a.corr) * env;      // same roundings as plain store + k_envelope
                                crow += a.ldc;
                            }
                        }
                        if (cc < c_hi) {
                            uint32_t v[16];
                            tmem_ld16_clipped(taddr, col0 + cc, v);
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                if (cc + j < c_hi && f_ok) *crow = __fmul_rn(__uint_as_float(v[j]), a.corr) * env;
                                crow += a.ldc;
                            }
                        }
                        if (a.nch > 1) {                                       // (TMEM loads are warp-wide: only the stores are predicated)
                            const float bf0 = __fmul_rn(tmem_ld1(taddr + col0), a.corr);
                            float t3[3] = {0.f, 0.f, 0.f};
                            const float e1[3] = {e1x, e1y, e1z};
#pragma unroll
                            for (int k = 0; k < 3; ++k) {
                                const int ch = ci + k;
                                const bool mine = ch >= c_lo && ch < c_hi;
                                if (mine || c_hi == a.nch) t3[k] = __fmul_rn(tmem_ld1(taddr + col0 + ch), a.corr);
                                if (mine && f_ok) cbase[(long)(col0 + ch) * a.ldc] = t3[k] * env + e1[k] * bf0;
                            }
                            if (c_hi == a.nch) {
                                const float vl = __fmul_rn(tmem_ld1(taddr + col0 + a.nch - 1), a.corr);
                                if (f_ok) cbase[(long)(col0 + a.nch - 1) * a.ldc] = vl * env + (el * bf0 + 2.f * (e1x * t3[0] + e1y * t3[1] + e1z * t3[2]));
                            }
                        }
                        continue;
                    }
                    // ---- epi 1: 16-column TMEM chunks aligned to the segment start; channel 0 is the value, nch - 1 the Laplacian.
                    // The addends of a chunk are fetched into registers ahead of the dependent chain; the shared-memory variant has a
                    // compile-time row stride (immediate offsets), the global one walks a pointer.
                    auto run_item = [&](auto smem_tag) {
                        constexpr bool SM = decltype(smem_tag)::value;
                        const float *arow = addend_rows(gi);
                        const long gstride = a.N_out;
                        float act_y = 0.f, act_d1 = 0.f, act_ssq = 0.f;
                        if (sg > 0) {                                           // later segments fetch the value column themselves
                            const float z0 = (__fmul_rn(tmem_ld1(taddr + col0), a.corr) + act_b) + (has_add ? arow[0] : 0.f);
                            act_y = tanhf(z0);
                            act_d1 = 1.f - act_y * act_y;
                        }
                        float *crow = cbase + (long)(col0 + c_lo) * a.ldc;
                        if (has_add) arow += SM ? (long)c_lo * 128 : (long)c_lo * gstride;
                        for (int cc = c_lo; cc < c_hi; cc += 16) {
                            uint32_t v[16];
                            const bool full = cc + 16 <= c_hi;
                            if (full) tmem_ld16(taddr + col0 + cc, v);
                            // addends: shared memory -> read in place (immediate offsets); global memory -> registers, ahead of the chain
                            float adg[SM ? 1 : 16];
                            if constexpr (!SM) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) adg[j] = (has_add && (full || cc + j < c_hi)) ? arow[j * gstride] : 0.f;
                            }
                            auto ad_at = [&](int j) -> float {
                                if constexpr (SM) return arow[j * 128];
                                else return adg[j];
                            };
#define ad_(j) ad_at(j)
                            if (full) tmem_ld_wait();
                            else tmem_ld16_clipped(taddr, col0 + cc, v);
                            if (cc == 0) {
                                const float z0 = (__fmul_rn(__uint_as_float(v[0]), a.corr) + act_b) + ad_(0);   // k_act's order: bias, then addend
                                act_y = tanhf(z0);
                                act_d1 = 1.f - act_y * act_y;
                            }
                            if (full && cc + 16 < a.nch && f_ok) {
                                // interior chunk: 16 tangent channels (or the value + 15 tangents), no bounds, no Laplacian column
                                {
                                    const float z = __fmul_rn(__uint_as_float(v[0]), a.corr) + ad_(0);
                                    float o = act_d1 * z;
                                    if (cc == 0) o = act_y; else act_ssq = fmaf(z, z, act_ssq);
                                    *crow = o;
                                    crow += a.ldc;
                                }
#pragma unroll
                                for (int j = 1; j < 16; ++j) {
                                    const float z = __fmul_rn(__uint_as_float(v[j]), a.corr) + ad_(j);
                                    act_ssq = fmaf(z, z, act_ssq);
                                    *crow = act_d1 * z;
                                    crow += a.ldc;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    const int cch = cc + j;
                                    if (cch < c_hi) {
                                        const float z = __fmul_rn(__uint_as_float(v[j]), a.corr) + ad_(j);
                                        float o = act_d1 * z;
                                        bool store = f_ok;
                                        if (cch == a.nch - 1) { o = o - 2.f * act_y * act_d1 * act_ssq; store = store && S == 1; }
                                        else act_ssq = fmaf(z, z, act_ssq);
                                        if (cch == 0) { o = act_y; act_ssq = 0.f; }
                                        if (store) *crow = o;
                                    }
                                    crow += a.ldc;
                                }
                            }
                            if (has_add) arow += SM ? 16 * 128 : 16 * gstride;
#undef ad_
                        }
                        if (S > 1) ssq_s[(gi * S + sg) * 128 + q * 32 + lane] = act_ssq;
                    };
                    if (add_smem) run_item(std::true_type{}); else run_item(std::false_type{});
                }
                if (S > 1 && a.epi == 1) {
                    // Laplacian rows: the partial sums of all segments of a group are complete after the quarter's barrier
                    asm volatile("bar.sync %0, 128;" ::"r"(8 + q) : "memory");
                    for (int gi = k4; gi < n_grp; gi += 4) {
                        const int col0 = gi * a.nch;
                        if (rows_valid - col0 < a.nch) continue;            // ragged last group of a segment: it has no Laplacian row here
                        const float *arow = addend_rows(gi);
                        const float z0 = (__fmul_rn(tmem_ld1(taddr + col0), a.corr) + act_b) + (has_add ? arow[0] : 0.f);
                        const float y = tanhf(z0), d1 = 1.f - y * y;
                        float ssq = 0.f;
                        for (int sg = 0; sg < S; ++sg) ssq += ssq_s[(gi * S + sg) * 128 + q * 32 + lane];
                        const float zl = __fmul_rn(tmem_ld1(taddr + col0 + a.nch - 1), a.corr) + (has_add ? arow[(long)(a.nch - 1) * astride] : 0.f);
                        if (f_ok) cbase[(long)(col0 + a.nch - 1) * a.ldc] = d1 * zl - 2.f * y * d1 * ssq;
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(8 + q) : "memory");      // the partial sums may be overwritten by the next tile
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster_relaxed(mapa_u32(&bar_tempty[buf], 0));      // this warp's share of the accumulator is drained
                    if (add_smem) mbar_arrive(&bar_aempty[slot]);
                }
                if (buf) tph1 ^= 1; else tph0 ^= 1;
                buf ^= 1;
            }
        } else {
        for (long t = pair; t < n_tiles; t += n_pairs, ++tcount) {
            const int ft = (int)(t % a.n_ft);
            const long rest = t / a.n_ft;
            const int rt = (int)(rest % a.n_rt), seg = (int)(rest / a.n_rt);
            const int m0 = rt * a.tile_rows;
            const int rows_valid = min(a.tile_rows, a.seg_len - m0);
            mbar_wait(&bar_tfull[buf], buf ? tph1 : tph0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
            uint32_t v[16];
            for (int c0 = g * 16; c0 < a.nmma; c0 += 32, ++chunk) {
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                if (c0 + 32 >= a.nmma) tc_fence_before();
                if (elected) tma_store_wait_read<1>();
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (elected && c0 + 32 >= a.nmma) mbar_arrive_cluster_relaxed(mapa_u32(&bar_tempty[buf], 0));    // the group's last chunk is in registers
                float *st = stage_base + (chunk & 1) * (TC_OUT_BYTES / 4) + q * 32 + lane;
#pragma unroll
                for (int j = 0; j < 16; ++j) st[j * 128] = __fmul_rn(__uint_as_float(v[j]), a.corr);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (elected) {
                    if (c0 < rows_valid) tma_store_3d(&map_c, stage_base + (chunk & 1) * (TC_OUT_BYTES / 4), ft * TC_FEAT + (int)rank * 128, m0 + c0, seg);
                    tma_store_commit();
                }
            }
            if (threadIdx.x == 6 * 32) TC_TL(3, tcount);
            if (buf) tph1 ^= 1; else tph0 ^= 1;
            buf ^= 1;
        }
        if (elected) tma_store_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                            // nobody frees TMEM while the peer's MMAs / loads may still touch it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// W[K][N] (row-major) -> Wt_hi / Wt_lo [N][K] (K-major), tf32-rounded halves
__global__ void k_split_transpose(const float *__restrict__ W, int K, int N, float *__restrict__ hi, float *__restrict__ lo) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= (long)K * N) return;
    int n = idx / K, k = idx - (long)n * K;
    float w = W[(long)k * N + n];
    float h = rna_tf32(w);
    hi[idx] = h;
    lo[idx] = rna_tf32(w - h);
}

// W[N][K] as stored -> tf32-rounded halves in the same layout (weights used transposed)
__global__ void k_split_plain(const float *__restrict__ W, long n, float *__restrict__ hi, float *__restrict__ lo) {
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float w = W[idx], h = rna_tf32(w);
    hi[idx] = h;
    lo[idx] = rna_tf32(w - h);
}

// ---------------------------------------------------------------------------------------- host side
// The tensor core adds every MMA (K = 8 tf32 products) to its FP32 accumulator with round-toward-zero.  Measured against
// fp64 on this part (tools/gemm_bias.py): sums of same-sign terms come out low by 7.9e-8 / 4.6e-7 / 1.07e-6 / 2.42e-6 / 2.85e-6
// (relative) for K = 16 / 64 / 128 / 256 / 320 -- three accumulations per K8 step, each losing half an ulp on average -- while the
// SPREAD of the error equals that of an FP32 FMA chain (2.3e-7).  Uncorrected, this bias is the whole difference between the
// tensor-core and the FP32 SIMT path (log psi^2 off by a systematic 2e-6 relative, 10x the FP32 floor).  The epilogues therefore
// scale every accumulator by 1 + bias(K): exact to first order for same-sign sums, and for cancelling sums the correction is as
// small as the result itself.
static float tc_rz_compensation(int K) {
    static const int kk[5] = {16, 64, 128, 256, 320};
    static const double bb[5] = {7.9e-8, 4.6e-7, 1.07e-6, 2.42e-6, 2.85e-6};
    static const bool off = getenv("DPE_TC_NO_RZ_COMP") != nullptr;
    if (off) return 1.0f;
    const int Kp = (K + TC_BK - 1) / TC_BK * TC_BK;          // zero-padded slabs still accumulate
    double b;
    if (Kp <= kk[0]) b = bb[0] * Kp / kk[0];
    else {
        int i = 1;
        while (i < 4 && Kp > kk[i]) ++i;
        b = bb[i - 1] + (bb[i] - bb[i - 1]) * (Kp - kk[i - 1]) / (kk[i] - kk[i - 1]);
    }
    return (float)(1.0 + b);
}

float tc_rz_comp(int K) { return tc_rz_compensation(K); }

struct TcWeight {
    const float *W;      // key: the [K, N] row-major weight the SIMT path would read
    int K, N;
    float *hi, *lo;      // [N, K]
    CUtensorMap map_hi, map_lo;
    CUtensorMap map_hi128, map_lo128;   // 128-row boxes for the CTA-pair kernel
    bool fresh;
    bool tr;             // W is stored [N, K]: the halves are a plain split (backward data products  C = A W^T)
};

struct TcState {
    std::vector<TcWeight> weights;
};

static int encode_w(CUtensorMap *map, float *ptr, int N, int K, int box_rows = TC_FEAT);
static int launch_gemm_tc_rows(dpe_model *m, const GemmArgs &g, const struct TcWeight *w, cudaStream_t s);

static int encode_w(CUtensorMap *map, float *ptr, int N, int K, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {TC_BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    return DPE_OK;
}

int tc_register_weight(dpe_model *m, const float *W, int K, int N, bool tr) {
    if (!m->tc) m->tc = new TcState();
    TcState *st = static_cast<TcState *>(m->tc);
    for (auto &c : st->weights)
        if (c.W == W && c.K == K && c.N == N && c.tr == tr) { c.fresh = false; return DPE_OK; }     // already registered (dpe_debug_gemm re-registers)
    if (tr && (K & 3)) return DPE_OK;       // optional use: a row pitch TMA cannot address leaves that product on the FP32 cores (the lookup finds nothing)
    TcWeight w;
    w.W = W; w.K = K; w.N = N; w.fresh = false; w.tr = tr;
    w.hi = w.lo = nullptr;
    DPE_CUDA(cudaMalloc(&w.hi, (size_t)K * N * sizeof(float)));
    if (cudaMalloc(&w.lo, (size_t)K * N * sizeof(float)) != cudaSuccess) { cudaFree(w.hi); return set_error(DPE_ERR_CUDA, "cudaMalloc of a split weight failed"); }
    int e;
    const int box_rows = N <= TR_N ? TR_N : TC_FEAT;      // narrow layers use the rows kernel (Wt resident in shared memory)
    e = encode_w(&w.map_hi, w.hi, N, K, box_rows);
    if (!e) e = encode_w(&w.map_lo, w.lo, N, K, box_rows);
    if (!e && N > TR_N) {
        e = encode_w(&w.map_hi128, w.hi, N, K, 128);
        if (!e) e = encode_w(&w.map_lo128, w.lo, N, K, 128);
    }
    if (e) { cudaFree(w.hi); cudaFree(w.lo); return e; }
    st->weights.push_back(w);
    return DPE_OK;
}

int tc_refresh_weights(dpe_model *m, cudaStream_t s) {
    if (!m->tc) return DPE_OK;
    TcState *st = static_cast<TcState *>(m->tc);
    for (auto &w : st->weights) {
        long n = (long)w.K * w.N;
        if (w.tr) k_split_plain<<<(int)((n + 255) / 256), 256, 0, s>>>(w.W, n, w.hi, w.lo);
        else k_split_transpose<<<(int)((n + 255) / 256), 256, 0, s>>>(w.W, w.K, w.N, w.hi, w.lo);
        DPE_LAUNCH_CHECK(m);
        w.fresh = true;
    }
    return DPE_OK;
}

void tc_destroy(dpe_model *m) {
    if (!m->tc) return;
    TcState *st = static_cast<TcState *>(m->tc);
    for (auto &w : st->weights) { cudaFree(w.hi); cudaFree(w.lo); }
    delete st;
    m->tc = nullptr;
}

int launch_gemm_tc(dpe_model *m, const GemmArgs &g, cudaStream_t s) {
    if (!m->tc) return DPE_ERR_UNSUPPORTED;
    TcState *st = static_cast<TcState *>(m->tc);
    const TcWeight *w = nullptr;
    for (auto &c : st->weights)
        if (c.W == g.W && c.K == g.K && c.N == g.N && c.tr == (g.w_tr != 0) && c.fresh) { w = &c; break; }
    if (!w) return DPE_ERR_UNSUPPORTED;
    if ((g.K & 3) || (g.lda & 3) || (reinterpret_cast<size_t>(g.A) & 15) || g.ldw != (g.w_tr ? g.K : g.N)) return DPE_ERR_UNSUPPORTED;
    if (g.N <= TR_N) return launch_gemm_tc_rows(m, g, w, s);
    // A and C must use the same segmentation (true for every caller in api.cu)
    if (g.a_seg_len != g.c_seg_len) return DPE_ERR_UNSUPPORTED;
    const int seg_len = g.a_seg_len < g.M ? g.a_seg_len : g.M;
    const int n_seg = g.M / seg_len;
    if ((long)n_seg * seg_len != g.M) return DPE_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return DPE_ERR_UNSUPPORTED;

    TcArgs a;
    a.C = g.C; a.ldc = g.ldc; a.c_seg_stride = n_seg > 1 ? g.c_seg_stride : 0; a.c_seg_off = g.c_seg_off; a.c_col_off = g.c_col_off;
    a.n_seg = n_seg; a.seg_len = seg_len; a.N_out = g.N; a.K = g.K;
    a.corr = tc_rz_compensation(g.K);
    static const bool pipe = getenv("DPE_TC_EPI_PIPE") != nullptr;
    a.pipe = pipe;
    a.spt = 1;
    a.w_seg_k = 0;
    a.epi = g.epi; a.nch = g.epi ? g.n_ch : 1;
    a.r = g.r; a.R = g.R; a.spa = g.spa; a.envw = g.envw; a.n_el = g.n_el; a.n_ion = g.n_ion; a.el_base = g.el_base;
    a.bias = g.bias; a.add = g.add; a.gpa = g.groups_per_add > 0 ? g.groups_per_add : 1;
    if (a.epi) {
        // group-aligned tiles: a tile owns whole (walker, electron) groups of C rows
        if (a.nch > 256 || seg_len % a.nch) return DPE_ERR_UNSUPPORTED;
        const int n_groups = seg_len / a.nch;
        int g_max = 256 / a.nch;
        const int pair_mode = tc_pair_mode();
        const int n_rt = (n_groups + g_max - 1) / g_max, gpt = (pair_mode > 1 && a.epi == 1) ? g_max : (n_groups + n_rt - 1) / n_rt;
        a.tile_rows = gpt * a.nch;
        a.nmma = (a.tile_rows + 15) / 16 * 16;
        a.n_rt = n_rt;
    } else if (n_seg > 1 && 2 * seg_len <= 256) {
        // short segments (forward pass: the spin block of one walker): pack several per tile with a 3-D TMA box
        a.spt = 256 / seg_len;
        if (a.spt > n_seg) a.spt = n_seg;
        a.tile_rows = seg_len * a.spt;
        a.nmma = (a.tile_rows + 15) / 16 * 16;
        a.n_rt = 1;
    } else {
        const int tiles_min = (seg_len + 255) / 256;
        a.nmma = (((seg_len + tiles_min - 1) / tiles_min) + 15) / 16 * 16;
        if (a.nmma > 256) a.nmma = 256;
        // few rows (the spin-mean term of a forward pass: one row per walker): narrower row tiles until about half the SMs have one -- the
        // tile's MMAs, not the weight loads, are what a lone CTA waits for (per-element arithmetic does not depend on the tile shape)
        static const bool no_small_tiles = getenv("DPE_TC_NO_SMALL_TILES") != nullptr;
        if (n_seg == 1 && !no_small_tiles) {
            const long n_ft0 = (g.N + TC_FEAT - 1) / TC_FEAT;
            while (a.nmma > 32 && ((seg_len + a.nmma - 1) / a.nmma) * n_ft0 * 2 <= m->n_sm) a.nmma = (a.nmma / 2 + 15) / 16 * 16;
        }
        a.tile_rows = a.nmma;
        a.n_rt = (seg_len + a.nmma - 1) / a.nmma;
    }
    a.n_ft = (g.N + TC_FEAT - 1) / TC_FEAT;

    CUtensorMap map_x;
    const long a_stride_rows = n_seg > 1 ? g.a_seg_stride : seg_len;
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)seg_len, (cuuint64_t)n_seg};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * sizeof(float), (cuuint64_t)a_stride_rows * g.lda * sizeof(float)};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)(a.spt > 1 ? seg_len : a.nmma), (cuuint32_t)a.spt};
    cuuint32_t es[3] = {1, 1, 1};
    void *base = const_cast<float *>(g.A + (long)g.a_seg_off * g.lda);
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed: %d", (int)r);

    // output map for the TMA-store epilogue: C viewed as (feature, row in segment, segment)
    CUtensorMap map_c = map_x;
    static const bool no_tma_store = getenv("DPE_TC_NO_TMA_STORE") != nullptr;
    a.tma_store = 0;
    bool have_map_c = false;
    if (!no_tma_store && a.epi <= 1 && a.spt == 1 && !(g.ldc & 3) && !(g.c_col_off & 3) && !(reinterpret_cast<size_t>(g.C) & 15)) {
        const long c_stride_rows = n_seg > 1 ? g.c_seg_stride : seg_len;
        cuuint64_t cdims[3] = {(cuuint64_t)g.N, (cuuint64_t)seg_len, (cuuint64_t)n_seg};
        cuuint64_t cstr[2] = {(cuuint64_t)g.ldc * sizeof(float), (cuuint64_t)c_stride_rows * g.ldc * sizeof(float)};
        cuuint32_t cbox[3] = {128, TC_OUT_ROWS, 1};
        cuuint32_t ces[3] = {1, 1, 1};
        void *cbase = g.C + (long)g.c_seg_off * g.ldc + g.c_col_off;
        CUresult rc = enc(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, cbase, cdims, cstr, cbox, ces, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc == CUDA_SUCCESS) { have_map_c = true; a.tma_store = a.epi == 0; }
    }

    if (int e = opt_in_smem(m, KID_GEMM_TC, k_gemm_tc_3xtf32)) return e;
    long n_tiles = (long)((a.n_seg + a.spt - 1) / a.spt) * a.n_rt * a.n_ft;
    int grid = (int)(n_tiles < m->n_sm ? n_tiles : m->n_sm);
    a.tl = nullptr;
    const char *tl_path = getenv("DPE_GEMM_TIMELINE");        // debug: role timeline of CTA 0 for launches with K >= 256 and M >= 1e6
    const bool tl_on = tl_path && g.K >= 256 && g.M >= 1000000;
    if (tl_on) {
        DPE_CUDA(cudaMalloc(&a.tl, TC_TL_TILES * 8 * sizeof(long long)));
        DPE_CUDA(cudaMemsetAsync(a.tl, 0, TC_TL_TILES * 8 * sizeof(long long), s));
    }
    bool launched_pair = false;
    const int use_pair = tc_pair_mode();
    // (the fused path must not depend on the batch size: chunked and single-pass evaluation have to agree bit for bit)
    bool pair_ok = use_pair && ((a.nmma + 31) & ~31) <= 256 && a.spt == 1 && (a.epi >= 1 || (a.epi == 0 && have_map_c && n_tiles >= m->n_sm));
    size_t smem2 = T2_SMEM_BYTES;
    a.n_st = T2_PLAIN_STAGES;
    CUtensorMap map_add = map_x;
    a.seg_split = 1;
    a.add_smem = 0;
    if (pair_ok && a.epi) {
        // fused epilogues: work items = channel segments of the (walker, electron) groups of a tile, four warps per TMEM lane quarter
        const int gpt = a.tile_rows / a.nch;
        static const int seg_env = getenv("DPE_TC_SEG_SPLIT") ? atoi(getenv("DPE_TC_SEG_SPLIT")) : 0;
        a.seg_split = seg_env > 0 ? seg_env : (gpt <= 1 ? 4 : (gpt <= 2 ? 2 : 1));
        if (a.nch < 8 * a.seg_split) a.seg_split = 1;
        if (a.nch < 8) pair_ok = false;             // forward-only passes (one channel per group): plain GEMM + the separate activation pass
        size_t part_bytes = a.epi == 1 && a.seg_split > 1 ? (size_t)gpt * a.seg_split * 512 : 0;
        // epi 1: the addend rows of a tile ([2 tiles][2 walkers][nch][128 features]) are TMA-prefetched into shared memory when a tile
        // touches at most two walkers and four K-stages still fit next to them (loads from global memory inside the dependent
        // epilogue chain cost the N2 layers 55 %); larger channel counts (3N + 2 > 48) read them from global memory
        static const bool no_add_smem = getenv("DPE_TC_ADD_GLOBAL") != nullptr;
        const size_t add_bytes = (size_t)4 * a.nch * 512;
        if (a.epi == 1 && g.add && !no_add_smem && gpt <= a.gpa && 4 * (size_t)T2_STAGE_BYTES + 1024 + add_bytes + part_bytes + 1024 <= 227 * 1024) {
            const long n_walkers = ((long)n_seg * seg_len / a.nch) / a.gpa;
            cuuint64_t adims[2] = {(cuuint64_t)g.N, (cuuint64_t)(n_walkers * a.nch)};
            cuuint64_t astr[1] = {(cuuint64_t)g.N * sizeof(float)};
            cuuint32_t abox[2] = {128, (cuuint32_t)a.nch};
            cuuint32_t aes[2] = {1, 1};
            CUresult ra = enc(&map_add, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(g.add), adims, astr, abox, aes,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (ra == CUDA_SUCCESS) { a.add_smem = 1; part_bytes += add_bytes; }
        }
        a.n_st = T2_STAGES;
        while (a.n_st > 3 && (size_t)a.n_st * T2_STAGE_BYTES + 1024 + part_bytes + 1024 > 227 * 1024) --a.n_st;
        smem2 = (size_t)a.n_st * T2_STAGE_BYTES + 1024 + part_bytes + 1024;
        if (smem2 > 227 * 1024 || (g.N & 3) || use_pair < 2) pair_ok = false;
    }
    // the fused epilogues are only worth running where they overlap the MMAs (double-buffered accumulators of the pair kernel)
    static const bool force_act = getenv("DPE_FUSE_ACT") != nullptr, force_env = getenv("DPE_FUSE_ENVELOPE") != nullptr;
    const bool force_fuse = a.epi == 1 ? force_act : force_env;
    if (a.epi && !pair_ok && !force_fuse) return DPE_ERR_UNSUPPORTED;
    if (pair_ok) {
        if (a.nmma & 31) a.nmma = (a.nmma + 31) & ~31;            // each CTA of the pair loads half of the MMA N rows
        CUtensorMap map_x2;
        cuuint32_t box2[3] = {TC_BK, (cuuint32_t)(a.nmma / 2), 1};
        CUresult r2 = enc(&map_x2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r2 == CUDA_SUCCESS) {
            const long pairs_wanted = n_tiles < m->n_sm / 2 ? n_tiles : m->n_sm / 2;
            const int grid2 = (int)pairs_wanted * 2;
            if (a.epi) {
                if (int e = opt_in_smem(m, KID_GEMM_TC2F, k_gemm_tc2_3xtf32<true>)) return e;
                k_gemm_tc2_3xtf32<true><<<grid2, (6 + T2_FUSED_EPI_WARPS) * 32, smem2, s>>>(map_x2, w->map_hi128, w->map_lo128, map_c, map_add, a);
            } else {
                if (int e = opt_in_smem(m, KID_GEMM_TC2P, k_gemm_tc2_3xtf32<false>)) return e;
                k_gemm_tc2_3xtf32<false><<<grid2, TC_THREADS, smem2, s>>>(map_x2, w->map_hi128, w->map_lo128, map_c, map_add, a);
            }
            launched_pair = true;
        } else if (a.epi && !force_fuse) {
            return DPE_ERR_UNSUPPORTED;
        }
    }
    if (!launched_pair) k_gemm_tc_3xtf32<<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(map_x, w->map_hi, w->map_lo, map_c, a);
    m->last_gemm_class = launched_pair ? 4 : 3;
    DPE_LAUNCH_CHECK(m);
    if (tl_on) {
        std::vector<long long> h(TC_TL_TILES * 8);
        DPE_CUDA(cudaStreamSynchronize(s));
        DPE_CUDA(cudaMemcpy(h.data(), a.tl, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.tl);
        if (FILE *f = fopen(tl_path, "a")) {
            fprintf(f, "# M=%d N=%d K=%d nmma=%d: tile mma_start mma_issued epi_start epi_end (clocks)\n", g.M, g.N, g.K, a.nmma);
            for (int t = 0; t < TC_TL_TILES; ++t)
                fprintf(f, "%d %lld %lld %lld %lld | %lld %lld %lld %lld\n", t, h[t * 8] - h[0], h[t * 8 + 1] - h[0], h[t * 8 + 2] - h[0], h[t * 8 + 3] - h[0], h[t * 8 + 4], h[t * 8 + 5], h[t * 8 + 6], h[t * 8 + 7]);
            fclose(f);
        }
    }
    return DPE_OK;
}

// Split-K product of the gradient / KFAC pass (grad.cu) on the CTA-pair kernel:
//   P[s][m, n] = sum_{k < Kc} At[m, s Kc + k] Bt[n, s Kc + k]        s < S = Rp / Kc
// At [S][Mt][Kc] FP32 (chunk-major; the reduction index -- walker x electron rows -- contiguous; split in shared memory like every X operand),
// Bt_hi / Bt_lo [Nb][Rp] tf32 halves (the role the layer weights play in a forward product).  A "segment" of the kernel is one K chunk: the
// X map walks chunks along its third dimension, the W maps get the chunk's K offset (w_seg_k), and segment s stores into P[s].
// The caller sums the S partial products in a fixed order (k_atb_reduce).
int launch_atb_tc(dpe_model *m, const float *At, float *Bt_hi, float *Bt_lo, long Rp, int Kc, int Mt, int Nb, float *part, cudaStream_t s) {
    static const bool off = getenv("DPE_ATB_TC") && atoi(getenv("DPE_ATB_TC")) == 0;
    if (off || m->gemm_path != 1 || tc_pair_mode() < 1) return DPE_ERR_UNSUPPORTED;
    if ((Nb & 3) || (Kc % TC_BK) || Rp % Kc || Rp >= (1L << 31) || (reinterpret_cast<size_t>(At) & 15) || (reinterpret_cast<size_t>(part) & 15)) return DPE_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return DPE_ERR_UNSUPPORTED;
    const int S = (int)(Rp / Kc);
    TcArgs a = {};
    a.C = part; a.ldc = Nb; a.c_seg_stride = Mt; a.c_seg_off = 0; a.c_col_off = 0;
    a.n_seg = S; a.seg_len = Mt; a.N_out = Nb; a.K = Kc;
    a.corr = tc_rz_compensation(Kc);
    a.spt = 1; a.w_seg_k = Kc; a.epi = 0; a.nch = 1; a.gpa = 1; a.seg_split = 1; a.n_st = T2_PLAIN_STAGES; a.tma_store = 1;
    const int tiles_min = (Mt + 255) / 256;
    a.nmma = (((Mt + tiles_min - 1) / tiles_min) + 31) / 32 * 32;     // each CTA of the pair loads half of the MMA N rows
    if (a.nmma > 256) a.nmma = 256;
    a.tile_rows = a.nmma;
    a.n_rt = (Mt + a.nmma - 1) / a.nmma;
    a.n_ft = (Nb + TC_FEAT - 1) / TC_FEAT;

    CUtensorMap map_x, map_c, map_h, map_l;
    cuuint32_t es[3] = {1, 1, 1};
    {
        cuuint64_t dims[3] = {(cuuint64_t)Kc, (cuuint64_t)Mt, (cuuint64_t)S};
        cuuint64_t strides[2] = {(cuuint64_t)Kc * sizeof(float), (cuuint64_t)Mt * Kc * sizeof(float)};      // chunk-major: At[s][m][k]
        cuuint32_t box[3] = {TC_BK, (cuuint32_t)(a.nmma / 2), 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(At), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(At) failed: %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Nb, (cuuint64_t)Mt, (cuuint64_t)S};
        cuuint64_t strides[2] = {(cuuint64_t)Nb * sizeof(float), (cuuint64_t)Mt * Nb * sizeof(float)};
        cuuint32_t box[3] = {128, TC_OUT_ROWS, 1};
        CUresult r = enc(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, part, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(P) failed: %d", (int)r);
    }
    if (int e = encode_w(&map_h, Bt_hi, Nb, (int)Rp, 128)) return e;
    if (int e = encode_w(&map_l, Bt_lo, Nb, (int)Rp, 128)) return e;
    if (int e = opt_in_smem(m, KID_GEMM_TC2P, k_gemm_tc2_3xtf32<false>)) return e;
    const long n_tiles = (long)S * a.n_rt * a.n_ft;
    const long pairs = n_tiles < m->n_sm / 2 ? n_tiles : m->n_sm / 2;
    k_gemm_tc2_3xtf32<false><<<(int)pairs * 2, TC_THREADS, T2_SMEM_BYTES, s>>>(map_x, map_h, map_l, map_c, map_x, a);
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

static int launch_gemm_tc_rows(dpe_model *m, const GemmArgs &g, const TcWeight *w, cudaStream_t s) {
    if (g.a_seg_len < g.M || g.c_seg_len < g.M || g.a_seg_off || g.c_seg_off) return DPE_ERR_UNSUPPORTED;   // plain rows only
    // forward pass (one channel per row group, no addend): bias + tanh in the epilogue.  (The same fusion in the wide kernels' staging epilogue
    // was measured slower than the separate k_act pass: 256 tanh per epilogue thread and tile are exposed behind ~10 k clocks of MMAs.)
    const bool act_fwd = g.epi == 1 && g.n_ch == 1 && !g.add && g.bias;
    if (g.K > TR_MAX_KB * TC_BK || (g.epi && !act_fwd)) return DPE_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return DPE_ERR_UNSUPPORTED;
    CUtensorMap map_x;
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.M, 1};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * sizeof(float), (cuuint64_t)g.M * g.lda * sizeof(float)};
    cuuint32_t box[3] = {TC_BK, TR_ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(g.A), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(X rows) failed: %d", (int)r);
    TrArgs a;
    a.C = g.C; a.ldc = g.ldc; a.c_col_off = g.c_col_off; a.M = g.M; a.N_out = g.N; a.K = g.K;
    a.corr = tc_rz_compensation(g.K);
    a.bias = act_fwd ? g.bias : nullptr;
    const int n_kb = (g.K + TC_BK - 1) / TC_BK;
    const size_t fixed = 2 * (size_t)n_kb * TR_WSLAB + 1024 + 1024;
    a.n_st = TR_STAGES;
    while (a.n_st > 2 && (size_t)a.n_st * TR_STAGE_BYTES + fixed > 227 * 1024) --a.n_st;
    const size_t smem_rows = (size_t)a.n_st * TR_STAGE_BYTES + fixed;
    if (int e = opt_in_smem(m, KID_GEMM_ROWS, k_gemm_tc_rows_3xtf32)) return e;
    long n_tiles = ((long)g.M + TR_ROWS - 1) / TR_ROWS;
    int grid = (int)(n_tiles < m->n_sm ? n_tiles : m->n_sm);
    k_gemm_tc_rows_3xtf32<<<grid, TR_THREADS, smem_rows, s>>>(map_x, w->map_hi, w->map_lo, a);
    m->last_gemm_class = 3;
    DPE_LAUNCH_CHECK(m);
    return DPE_OK;
}

}  // namespace dpe
