// Tangent / Laplacian stage of the determinants on the tensor cores (forward-Laplacian of log|det A|).
//
// For every determinant A (N x N, one per walker and per det index) and every forward-Laplacian channel c the energy needs
//     g_c   = tr(Ainv dA_c)             (c = 1 .. 3N: tangents,  c = C-1: the Laplacian channel)
//     t2_c  = tr((Ainv dA_c)^2)         (tangents only)
// which the reference obtains by differentiating jnp.linalg.slogdet (model/orbitals/__init__ -> slogdet inside
// wavefunction.py:228-259 under folx's forward_laplacian, hamiltonian.py:197-213).  Both traces are invariant under the
// similarity P_c = Ainv dA_c  ->  Q_c = dA_c Ainv, and Q_c contracts over the ORBITAL index, which is the contiguous
// index of the orbital tensor mo[b][i][c][det*N + q] the envelope kernel wrote.  So the whole stage is one batched GEMM
//     Q[(c, i), i'] = sum_q mo[b][i][c][det N + q] * Ainv[q][i']
// with the rows (c, i) of one determinant streamed straight out of `mo` by a 4-D TMA box -- no gather, no transpose --
// and Ainv^T (FP64 Gauss-Jordan in orbitals_det.cu, tf32-split, zero padded to NP = roundup16(N)) resident in shared memory.
// A TMA box must start on a 16-byte boundary, so the slab starts at the aligned column below det*N and Ainv^T is stored
// shifted by (det*N) mod 4; its zero padding cancels whatever the 16-float K slabs pick up outside the determinant.
//
// Kernel structure = the rows GEMM of gemm_tc.cu: warp 0 TMA producer, warp 1 tcgen05.mma issuer (3xTF32: hi*lo + lo*hi +
// hi*hi, FP32 accumulate in TMEM), warps 2-5 split the raw FP32 slab into tf32 hi / lo in place, warps 6-13 drain the
// double-buffered accumulators (one accumulator row per thread).  The epilogue never writes Q to global memory: it parks the 256 x NP tile in shared
// memory, takes the diagonal and the 2 x 2 principal minors of each N x N block (g_c^2 - t2_c = 2 e2(Q_c); the minor
// form keeps the cancellation of the two sums for an ill-conditioned A inside each minor) and emits the determinant record
// [logdet, sign, lap', g_1 .. g_3N] that k_combine consumes.
#include "dpe_internal.cuh"
#include "tc_common.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace dpe {

constexpr int DT_ROWS = 256;                     // rows per tile (2 x M = 128)
constexpr int DT_X_BYTES = DT_ROWS * TC_ROWB;    // 16 KB raw -> hi, + 16 KB lo
constexpr int DT_STAGE_BYTES = 2 * DT_X_BYTES;
constexpr int DT_MAX_STAGES = 4;
constexpr int DT_EPI_WARPS = 8;                  // one accumulator row per epilogue thread
constexpr int DT_THREADS = (6 + DT_EPI_WARPS) * 32;      // per epilogue group: the G = 2 instantiation runs (6 + 2 * DT_EPI_WARPS) warps
constexpr int DT_TMEM_COLS = 512;                // 2 buffers x 2 halves x 2 NP (NP <= 64)

struct DtArgs {
    float *det;            // [unit][rec]
    int N, NP, C, n_det, cpt, n_tiles, n_kb, n_stages, n_half, n_qbuf, qs, rec;
    long n_units;
    long long *tl;         // debug timeline (DPE_DET_TIMELINE): [tile][8] clock64 stamps of CTA 0, or nullptr
};

constexpr int DT_TL_TILES = 256;
#define DT_TL(role, idx) do { if (a.tl && blockIdx.x == 0 && (idx) < DT_TL_TILES) a.tl[(idx) * 12 + (role)] = clock64(); } while (0)

static __device__ __forceinline__ double dt_warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 2 x 2 minor a d - b c with one rounding: p = rn(b c), e = b c - p exactly, and fma(a, d, -p) rounds a d - p once.
// (The first version of this epilogue accumulated tr(Q^2) and g^2 separately in FP64: F2F.F64.F32 runs on the XU pipe at
// about one result per clock per SM here -- ncu showed XU at 124 % -- and a float-float TwoSum version was latency bound.)
static __device__ __forceinline__ float minor2(float a, float d, float b, float c) {
    const float p = __fmul_rn(b, c);
    const float e = __fmaf_rn(b, c, -p);
    return __fsub_rn(__fmaf_rn(a, d, -p), e);
}
static __device__ __forceinline__ void dt_epi_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(DT_EPI_WARPS * 32) : "memory"); }

// G = number of epilogue groups (8 warps each).  The epilogue is a latency chain per thread (TMEM -> shared memory, then the walk over
// the 2 x 2 minors), so with G = 2 two groups take alternate tiles -- group g always drains accumulator g -- each with its own Q
// buffer and named barrier; the per-determinant sums of the two groups meet in shared memory (two addends: order-independent).
template <int G>
__global__ void __launch_bounds__((6 + G * DT_EPI_WARPS) * 32, 1)
k_det_trace_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
               const __grid_constant__ CUtensorMap map_wl, DtArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
    const int wslab = a.NP * TC_ROWB;                       // one K slab of Ainv^T (hi or lo)
    const int whalf = a.n_kb * wslab;                       // per buffer: n_kb x [hi slab | lo slab]
    uint8_t *wbase = smem + a.n_stages * DT_STAGE_BYTES;    // [2 buffers][hi | lo]
    const int QS = a.qs;      // row stride = 2 mod 4: the cyclic row / column / diagonal walks of the epilogue all advance by QS + 1
                              // words from lane to lane (odd: conflict free); the row stores see a 2-way conflict
    float *Qs = reinterpret_cast<float *>(wbase + 4 * whalf);   // [n_qbuf][256][QS]
    double *red = reinterpret_cast<double *>(Qs + a.n_qbuf * DT_ROWS * QS);     // [2 groups][2][8]; (NP + 1) * 1024 + 1024 bytes after W keeps it 8-byte aligned
    double *unit_acc = red + 4 * DT_EPI_WARPS;                                  // [4] per-determinant sums where both groups contribute
    int *unit_cnt = reinterpret_cast<int *>(unit_acc + 4);                      // [4]
    uint64_t *bars = reinterpret_cast<uint64_t *>(unit_cnt + 4);
    uint64_t *bar_full = bars, *bar_split = bars + DT_MAX_STAGES, *bar_empty = bars + 2 * DT_MAX_STAGES;
    uint64_t *bar_tfull = bars + 3 * DT_MAX_STAGES, *bar_tempty = bar_tfull + 2;
    uint64_t *bar_wfull = bar_tempty + 2, *bar_wempty = bar_wfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_wempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = a.N, NP = a.NP, C = a.C;
    const int NA = (2 * NP + 31) & ~31;                           // TMEM columns per accumulator [hi hi + lo hi | hi lo], 32-column aligned
    const uint32_t xbytes = (uint32_t)(TC_ROWB * N * a.cpt);      // one TMA box: cpt channels x N electrons x 64 B

    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) { unit_acc[k] = 0.0; unit_cnt[k] = 0; }
        for (int s = 0; s < DT_MAX_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_split[s], 128); mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], DT_EPI_WARPS * 32);
            mbar_init(&bar_wfull[b], 1); mbar_init(&bar_wempty[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(DT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int ui = 0, tc_ = 0;
            for (long u = blockIdx.x; u < a.n_units; u += gridDim.x, ++ui) {
                const int wb = ui & 1;
                const long b = u / a.n_det;
                const int dt = (int)(u - b * a.n_det);
                mbar_wait(&bar_wempty[wb], (((uint32_t)ui >> 1) & 1u) ^ 1u);
                mbar_expect_tx(&bar_wfull[wb], 2 * whalf);
                uint8_t *w0 = wbase + wb * 2 * whalf;
                for (int kb = 0; kb < a.n_kb; ++kb) {        // hi and lo of one K slab sit back to back: one B operand of 2 NP rows
                    tma_load_2d(w0 + (2 * kb) * wslab, &map_wh, &bar_wfull[wb], kb * TC_BK, (int)(u * NP));
                    tma_load_2d(w0 + (2 * kb + 1) * wslab, &map_wl, &bar_wfull[wb], kb * TC_BK, (int)(u * NP));
                }
                for (int t = 0; t < a.n_tiles; ++t, ++tc_)
                    for (int kb = 0; kb < a.n_kb; ++kb) {
                        mbar_wait(&bar_empty[stage], phase ^ 1);
                        if (kb == 0) DT_TL(0, tc_);
                        mbar_expect_tx(&bar_full[stage], xbytes);
                        tma_load_4d(smem + stage * DT_STAGE_BYTES, &map_x, &bar_full[stage], ((dt * N) & ~3) + kb * TC_BK, 0, 1 + t * a.cpt, (int)b);
                        if (++stage == a.n_stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // two MMAs per K8 step and row half:  xh * [wh | wl] -> columns [0, 2 NP) = [hi hi | hi lo],  xl * wh accumulated onto
            // columns [0, NP).  X_hi is read from shared memory once for both products; the epilogue adds the two column groups.
            const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * NP) >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0, tph0 = 0, tph1 = 0;
            int buf = 0, ui = 0, tc_ = 0;
            for (long u = blockIdx.x; u < a.n_units; u += gridDim.x, ++ui) {
                const int wb = ui & 1;
                mbar_wait(&bar_wfull[wb], ((uint32_t)ui >> 1) & 1u);
                const uint32_t w0 = smem_u32(wbase + wb * 2 * whalf);
                for (int t = 0; t < a.n_tiles; ++t, ++tc_) {
                    mbar_wait(&bar_tempty[buf], (buf ? tph1 : tph0) ^ 1);
                    tc_fence_after();
                    DT_TL(3, tc_);
                    for (int kb = 0; kb < a.n_kb; ++kb) {
                        mbar_wait(&bar_split[stage], phase);
                        tc_fence_after();
                        if (kb == 0) DT_TL(4, tc_);
                        const uint32_t xh = smem_u32(smem + stage * DT_STAGE_BYTES), xl = xh + DT_X_BYTES;
                        const uint32_t wh = w0 + 2 * kb * wslab;
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 8; ++kk) {
                            const uint32_t ko = kk * 32;
                            const uint64_t dw = make_desc_sw64(wh + ko);
                            for (int h = 0; h < a.n_half; ++h) {
                                const uint32_t d = tmem_base + (buf * 2 + h) * NA;
                                tc_mma_tf32(d, make_desc_sw64(xh + h * 128 * TC_ROWB + ko), dw, idesc2, (kb | kk) ? 1u : 0u);
                                tc_mma_tf32(d, make_desc_sw64(xl + h * 128 * TC_ROWB + ko), dw, idesc1, 1u);
                            }
                        }
                        tc_commit(&bar_empty[stage]);
                        if (++stage == a.n_stages) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(&bar_tfull[buf]);
                    DT_TL(5, tc_);
                    if (buf) tph1 ^= 1; else tph0 ^= 1;
                    buf ^= 1;
                }
                tc_commit(&bar_wempty[wb]);       // fires once every MMA that read this copy of Ainv^T has retired
            }
        }
    } else if (warp < 6) {
        const int tid = threadIdx.x - 64;
        const int n_vec = (int)(xbytes / 16);
        int stage = 0, sc_ = 0; uint32_t phase = 0;
        for (long u = blockIdx.x; u < a.n_units; u += gridDim.x)
            for (int it = 0; it < a.n_tiles * a.n_kb; ++it, ++sc_) {
                mbar_wait(&bar_full[stage], phase);
                if (tid == 0) DT_TL(1, sc_ / a.n_kb);
                float4 *xh = reinterpret_cast<float4 *>(smem + stage * DT_STAGE_BYTES);
                float4 *xl = reinterpret_cast<float4 *>(smem + stage * DT_STAGE_BYTES + DT_X_BYTES);
                for (int base = tid; base < n_vec; base += 4 * 128) {     // loads batched ahead of the stores (they may alias for the compiler)
                    float4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (base + k * 128 < n_vec) v[k] = xh[base + k * 128];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (base + k * 128 < n_vec) {
                            float4 h, l;
                            h.x = rna_tf32(v[k].x); h.y = rna_tf32(v[k].y); h.z = rna_tf32(v[k].z); h.w = rna_tf32(v[k].w);
                            l.x = rna_tf32(v[k].x - h.x); l.y = rna_tf32(v[k].y - h.y); l.z = rna_tf32(v[k].z - h.z); l.w = rna_tf32(v[k].w - h.w);
                            xh[base + k * 128] = h;
                            xl[base + k * 128] = l;
                        }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&bar_split[stage]);
                if (tid == 0) DT_TL(2, sc_ / a.n_kb);
                if (++stage == a.n_stages) { stage = 0; phase ^= 1; }
            }
    } else {
        const int grp = (warp - 6) / DT_EPI_WARPS, gw = (warp - 6) - grp * DT_EPI_WARPS;     // epilogue group, warp within the group
        const int q = warp & 3, hh = gw >> 2;               // TMEM lane quarter (= warp % 4) and accumulator half of this warp
        const int mrow = hh * 128 + q * 32 + lane;          // the tile row this thread owns
        const int cl = mrow / N, i = mrow - cl * N;         // (channel within the tile, electron) of that row
        // cyclic partners (i + 1 .. i + n_pairs) mod N: every unordered pair once; for even N the antipodal pair goes to the lower index
        const int n_pairs = ((N - 1) >> 1) + ((!(N & 1) && 2 * i < N) ? 1 : 0);
        double *red_g = red + grp * 2 * DT_EPI_WARPS;
        const bool both = G == 2 && a.n_tiles >= 2;         // both groups see tiles of every determinant
        int ui = 0, tcount = 0;
        float *pend = nullptr;                              // record whose lap' still waits for the group-wide sum (thread mrow == 0)
        int pend_par = 0, pend_ui = 0;
        auto finish_pending = [&]() {
            const double *rd = red_g + pend_par * DT_EPI_WARPS;
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < DT_EPI_WARPS; ++w) tot += rd[w];
            if (!both) {
                pend[2] = (float)tot;
            } else {
                const int slot = pend_ui & 3;
                atomicAdd(&unit_acc[slot], tot);
                __threadfence_block();
                if (atomicAdd(&unit_cnt[slot], 1) == 1) {   // the other group's share is in: a + b, whichever came first
                    __threadfence_block();
                    pend[2] = (float)atomicAdd(&unit_acc[slot], 0.0);
                    unit_acc[slot] = 0.0;
                    unit_cnt[slot] = 0;
                }
            }
            pend = nullptr;
        };
        for (long u = blockIdx.x; u < a.n_units; u += gridDim.x, ++ui) {
            float *out = a.det + u * a.rec;
            // per determinant:  lap' = tr(Ainv lapA) + sum_k (g_k^2 - tr(Q_k^2)) = tr(Ainv lapA) + 2 sum_k e2(Q_k), with e2 the sum
            // of the 2 x 2 principal minors.  g_k^2 and tr(Q_k^2) cancel to many digits for an ill-conditioned A (Q_k close to
            // rank one); in the minor form that cancellation happens inside each minor, where minor2() resolves it exactly.
            float e2 = 0.f, lap_tr = 0.f;
            int n_mine = 0;
            for (int t = 0; t < a.n_tiles; ++t, ++tcount) {
                const int buf = tcount & 1;
                if (G == 2 && buf != grp) continue;            // the other group's tile
                ++n_mine;
                const uint32_t tph = ((uint32_t)tcount >> 1) & 1u;
                const int c0 = 1 + t * a.cpt;
                const int nc = min(a.cpt, C - c0);
                const int rows = nc * N;
                const bool one_q = G == 2 || a.n_qbuf == 1;    // this group has a single Q buffer
                float *Q = Qs + (G == 2 ? grp : (a.n_qbuf == 2 ? (tcount & 1) : 0)) * DT_ROWS * QS;
                mbar_wait(&bar_tfull[buf], tph);
                tc_fence_after();
                if (mrow == 0 && grp == 0) DT_TL(6, tcount);
                if (one_q) dt_epi_sync(grp);                   // single Q buffer: everybody is done reading the previous tile
                if (hh < a.n_half) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + hh) * NA;
                    float2 *d0 = reinterpret_cast<float2 *>(Q + mrow * QS);
                    for (int ch = 0; ch < N; ch += 16) {
                        uint32_t v0[16], v1[16];
                        tmem_ld16(taddr + ch, v0);
                        tmem_ld16(taddr + NP + ch, v1);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            if (ch + j < N)
                                d0[(ch + j) >> 1] = make_float2(__fadd_rn(__uint_as_float(v0[j]), __uint_as_float(v1[j])),
                                                                __fadd_rn(__uint_as_float(v0[j + 1]), __uint_as_float(v1[j + 1])));
                    }
                }
                tc_fence_before();
                mbar_arrive(&bar_tempty[buf]);
                if (mrow == N - 1 && grp == 0) DT_TL(8, tcount);
                dt_epi_sync(grp);                              // Q tile complete (and, with two buffers, the tile before last retired)
                if (mrow == 0 && pend) finish_pending();
                if (mrow == N - 1 && grp == 0) DT_TL(9, tcount);
                if (mrow < rows && c0 + cl < C - 1) {          // tangent channel: this row's share of the 2 x 2 principal minors of Q_c
                    const float *__restrict__ row = Q + mrow * QS;
                    const float *__restrict__ blk = Q + cl * N * QS;
                    const float qii = row[i];
                    constexpr int PB = G == 2 ? 4 : 8;                 // operands of up to PB pairs are fetched ahead of the arithmetic;
                    for (int d0 = 1; d0 <= n_pairs; d0 += PB) {        // no branches: pairs past the end read row i itself and add 0
                        float pd[PB], pb[PB], pc[PB];
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            int o = i + d0 + k;
                            o = o >= N ? o - N : o;
                            o = d0 + k <= n_pairs ? o : i;
                            pd[k] = blk[o * QS + o]; pb[k] = row[o]; pc[k] = blk[o * QS + i];
                        }
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const float mm = minor2(qii, pd[k], pb[k], pc[k]);
                            e2 = __fadd_rn(e2, d0 + k <= n_pairs ? mm : 0.f);
                        }
                    }
                }
                if (mrow == N - 1 && grp == 0) DT_TL(10, tcount);
                {
                    // g_c = tr(dA_c Ainv): one thread per channel sums the diagonal of its block -- in a warp whose rows the tile does
                    // not use, when there is one (it then runs beside the minors instead of after them), else the block's last row
                    const int spare_base = (rows + 31) & ~31;
                    const bool spare = spare_base + nc <= DT_ROWS;
                    const int cd = spare ? mrow - spare_base : cl;
                    if (spare ? (cd >= 0 && cd < nc) : (mrow < rows && i == N - 1)) {
                        const float *__restrict__ blk = Q + cd * N * QS;
                        float g0 = 0.f, g1 = 0.f;
                        int o = 0;
#pragma unroll 4
                        for (; o + 1 < N; o += 2) { g0 = __fadd_rn(g0, blk[o * QS + o]); g1 = __fadd_rn(g1, blk[(o + 1) * QS + o + 1]); }
                        if (o < N) g0 = __fadd_rn(g0, blk[o * QS + o]);
                        const float gk = __fadd_rn(g0, g1);
                        if (c0 + cd < C - 1) out[3 + c0 + cd - 1] = gk;
                        else lap_tr = gk;                      // tr(Ainv lapA)
                    }
                }
                if (mrow == N - 1 && grp == 0) DT_TL(7, tcount);
            }
            if (n_mine) {
                double acc = 2.0 * (double)e2 + (double)lap_tr;
                acc = dt_warp_sum(acc);
                if (lane == 0) red_g[(ui & 1) * DT_EPI_WARPS + gw] = acc;
                if (mrow == 0 && pend) {                      // two determinants in a row without a barrier of this group in between cannot happen:
                    /* unreachable: every determinant this group takes part in has at least one tile with two barriers */
                }
                pend = out;                                   // summed by thread mrow == 0 after the group's next barrier
                pend_par = ui & 1;
                pend_ui = ui;
            }
        }
        dt_epi_sync(grp);
        if (mrow == 0 && pend) finish_pending();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(DT_TMEM_COLS));
    }
}

int launch_det_trace_tc(dpe_model *m, int Bc, int C, const float *mo, const float *ainv_hi, const float *ainv_lo, int NP,
                        float *det, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const int N = d.n_el, cols = d.n_dets * N;
    if (N < 1 || N > 64 || NP > 64 || C < 3 || (cols & 3)) return DPE_ERR_UNSUPPORTED;     // TMA strides are 16-byte multiples
    EncodeTiledFn enc = get_encode();
    if (!enc) return DPE_ERR_UNSUPPORTED;

    DtArgs a;
    a.det = det; a.N = N; a.NP = NP; a.C = C; a.n_det = d.n_dets; a.rec = C - 2 + 3;
    a.n_units = (long)Bc * d.n_dets;
    const int cpt_max = std::min(DT_ROWS / N, C - 1);
    a.n_tiles = (C - 1 + cpt_max - 1) / cpt_max;
    a.cpt = (C - 1 + a.n_tiles - 1) / a.n_tiles;          // balanced channel tiles: c = 1 .. C-1
    a.n_kb = NP / TC_BK;
    a.n_half = N * a.cpt > 128 ? 2 : 1;

    CUtensorMap map_x, map_wh, map_wl;
    {   // mo[b][i][c][col] viewed as (col, i, c, b); one box = 16 columns x N electrons x cpt channels of one walker
        cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)Bc};
        cuuint64_t strides[3] = {(cuuint64_t)C * cols * sizeof(float), (cuuint64_t)cols * sizeof(float), (cuuint64_t)N * C * cols * sizeof(float)};
        cuuint32_t box[4] = {TC_BK, (cuuint32_t)N, (cuuint32_t)a.cpt, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(mo), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(mo) failed: %d", (int)r);
    }
    for (int hl = 0; hl < 2; ++hl) {
        cuuint64_t dims[2] = {(cuuint64_t)NP, (cuuint64_t)a.n_units * NP};
        cuuint64_t strides[1] = {(cuuint64_t)NP * sizeof(float)};
        cuuint32_t box[2] = {TC_BK, (cuuint32_t)NP};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(hl ? &map_wl : &map_wh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(hl ? ainv_lo : ainv_hi), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return set_error(DPE_ERR_CUDA, "cuTensorMapEncodeTiled(Ainv) failed: %d", (int)r);
    }
    a.qs = ((N + 1) & ~3) + 2;                            // smallest stride >= N with stride = 2 (mod 4)
    // shared-memory plan: as many TMA stages as fit (HBM latency), two Q buffers if possible (one barrier per tile)
    const size_t w_bytes = 4 * (size_t)a.n_kb * NP * TC_ROWB, q_bytes = (size_t)DT_ROWS * a.qs * sizeof(float);
    const size_t fixed = w_bytes + (4 * DT_EPI_WARPS + 4) * sizeof(double) + 4 * sizeof(int) + (3 * DT_MAX_STAGES + 8) * sizeof(uint64_t) + 16 + 1024;
    const int options[5][2] = {{4, 2}, {3, 2}, {3, 1}, {2, 2}, {2, 1}};
    size_t smem = 0;
    for (const auto &o : options) {
        smem = fixed + (size_t)o[0] * DT_STAGE_BYTES + o[1] * q_bytes;
        a.n_stages = o[0]; a.n_qbuf = o[1];
        if (smem <= 227 * 1024) break;
    }
    if (smem > 227 * 1024) return DPE_ERR_UNSUPPORTED;
    // two epilogue groups need a Q buffer each (the n_qbuf = 2 plans)
    static const int groups_env = getenv("DPE_DET_EPI_GROUPS") ? atoi(getenv("DPE_DET_EPI_GROUPS")) : 2;
    // measured: benzene (22 tiles per determinant) 71.6 -> 68.0 ms with two groups, N2 (3 tiles) unchanged -- the stage is bound by the depth of
    // its TMA pipeline (three 32 KB stages per tile next to the Q buffers), not by the epilogue alone
    const int n_groups = (groups_env == 2 && a.n_qbuf == 2 && a.n_tiles >= 8) ? 2 : 1;
    if (n_groups == 2) { if (int e = opt_in_smem(m, KID_DET_TRACE, k_det_trace_tc<2>)) return e; }
    else if (int e = opt_in_smem(m, KID_GRAD_B, k_det_trace_tc<1>)) return e;
    const int grid = (int)(a.n_units < m->n_sm ? a.n_units : m->n_sm);
    a.tl = nullptr;
    const char *tl_path = getenv("DPE_DET_TIMELINE");          // debug: dump the role timeline of CTA 0 (clock64 stamps)
    if (tl_path) {
        DPE_CUDA(cudaMalloc(&a.tl, DT_TL_TILES * 12 * sizeof(long long)));
        DPE_CUDA(cudaMemsetAsync(a.tl, 0, DT_TL_TILES * 12 * sizeof(long long), s));
    }
    if (n_groups == 2) k_det_trace_tc<2><<<grid, (6 + 2 * DT_EPI_WARPS) * 32, smem, s>>>(map_x, map_wh, map_wl, a);
    else k_det_trace_tc<1><<<grid, DT_THREADS, smem, s>>>(map_x, map_wh, map_wl, a);
    DPE_LAUNCH_CHECK(m);
    if (tl_path) {
        std::vector<long long> h(DT_TL_TILES * 12);
        DPE_CUDA(cudaStreamSynchronize(s));
        DPE_CUDA(cudaMemcpy(h.data(), a.tl, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.tl);
        if (FILE *f = fopen(tl_path, "w")) {
            fprintf(f, "# tile tma_issue split_begin split_end mma_tempty mma_split mma_commit epi_tfull epi_end epi_sts_done epi_bar_done epi_pairs_done -  (clocks since first stamp); n_tiles/unit=%d n_kb=%d stages=%d qbuf=%d\n", a.n_tiles, a.n_kb, a.n_stages, a.n_qbuf);
            long long t0 = h[0];
            for (int t = 0; t < DT_TL_TILES; ++t) {
                fprintf(f, "%d", t);
                for (int r = 0; r < 12; ++r) fprintf(f, " %lld", h[t * 12 + r] ? h[t * 12 + r] - t0 : -1);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return DPE_OK;
}

}  // namespace dpe
