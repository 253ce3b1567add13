// Internal declarations shared by the kernels of libdpe_b200.so (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include "../../include/dpe_b200.h"

#define DPE_MAX_LEAVES (1 + DPE_MAX_ITER * 16 + 6)
#define DPE_GEOM_SLOTS 8

namespace dpe {

// Channel convention of every "tangent-carrying" tensor: channel 0 = value, channels 1..n_t = d/dx_k,
// channel n_t+1 = Laplacian.  Forward-only mode has a single channel (n_t = -1 -> C = 1).
struct Leaf { int64_t off; int64_t size; int32_t rows, cols; };

struct Dense { const float *w, *b; int din, dout; };

struct IterParams {
    Dense w_same, w_diff, h_map, h_ion_map, h_el, h_same, h_diff, h_el_ion;  // views into the flat params
    float *w_main;  // [k_main, d_out]  rows: h_one (d_in) | conv_ee (emb) | conv_eI (dE)  of h_el.w
    float *w_mean;  // [2*d_in, d_out]  rows: mean_up | mean_dn                            of h_el.w
    float *him;     // [n_ion, dE] = tanh(Lin_h_ion_map(h_ion[Z]))   (geometry-only)
    int d_in, dP, dE, d_out, k_main;
    int pair_next, eion_next;  // widths after this iteration's pair / el-ion layers (0 on the last iteration)
};

struct GemmArgs {
    const float *A; int lda;
    int a_seg_len, a_seg_stride, a_seg_off;   // row m -> (m / seg_len) * seg_stride + seg_off + m % seg_len
    const float *W; int ldw;                  // [K, N] row-major
    int w_tr = 0;                             // 1: W is stored [N, K] row-major (C = A W^T with a layer's own weight: backward data products, tensor-core path only)
    float *C; int ldc;
    int c_seg_len, c_seg_stride, c_seg_off, c_col_off;
    int M, N, K;
    // optional fused epilogue (tensor-core path only; the SIMT path ignores it and the caller runs the separate kernel)
    int epi = 0, n_ch = 1;
    const float *r = nullptr, *R = nullptr, *spa = nullptr, *envw = nullptr;
    int n_el = 0, n_ion = 0, el_base = 0;
    // epi == 1: tanh rule of the dense layer (mlp.py:45-69) on bias + per-walker addend, see k_act
    const float *bias = nullptr, *add = nullptr;
    int groups_per_add = 1;
};

// Workspace layout for one chunk of Bc walkers with C channels (byte offsets).
struct WsLayout {
    size_t x[2], hm, mean, add, pw, ei, mo, det, ainv, tao_g, epot, lp, total_chunk;
    size_t ei_it[DPE_MAX_ITER];   // offsets (bytes) of the per-iteration el-ion convolution blocks
    size_t pw_it[DPE_MAX_ITER];
    // per-call (full batch) scratch for the Metropolis step
    size_t r_prop, lp_prop, thr, new_keys, log_q, mask, ctrl, total_mcmc;
    int ldx;       // row stride (floats) of the X buffers
};

struct McmcGraphCache;
}  // namespace dpe

struct dpe_model {
    dpe_dims dims;
    int n_leaves;
    dpe::Leaf leaves[DPE_MAX_LEAVES];
    int64_t n_params;
    float *params;      // device copy of the flat vector
    float *derived;     // device: w_main / w_mean / softplus(alpha) / him blocks
    size_t derived_floats;
    dpe::IterParams it[DPE_MAX_ITER];
    const float *h_ion_emb;                       // [V, F]
    const float *bf_w[2], *alpha[2], *env_w[2];   // up, dn
    float *sp_alpha[2];                           // softplus(alpha) [n_ion, n_det*n_el]
    float *tao_w;                                 // TAO: [d_last, n_ion * n_det * n_el] backflow matrix, column = (ion, det, orbital)
    float *tao_ex[2];                             // TAO: exponents [n_ion, n_det * n_el] for same-spin / different-spin (electron, orbital)
    bool tao_set;
    float *R_dev;  // [n_ion,3]
    float *Z_dev;  // [n_ion] as float
    float *eii_dev;                               // [1] ion-ion repulsion (hamiltonian.py:25-31), float32 like the reference
    int32_t *geom_flag_dev;                       // [1] != 0 after a device-side set_geometry saw a Z outside [z_min, z_max]
    // pinned staging ring of the host-array set_geometry (asynchronous H2D without a stream synchronisation)
    float *geom_pin;                              // [DPE_GEOM_SLOTS][4 * n_ion + 1]
    cudaEvent_t geom_ev[DPE_GEOM_SLOTS];
    int geom_slot;
    int det_flags;                                // dpe_set_det_path: bit 0 generic kernel also for N <= 16, bit 1 SIMT tangent stage
    unsigned long long smem_opted;                // bit k: kernel k of this device was opted in to the full dynamic shared memory
    bool params_set, geom_set;
    int gemm_path;
    int64_t launches;
    dpe::McmcGraphCache *mcmc_graphs;           // captured Metropolis step sequences (api.cu), nullptr until first use
    int mcmc_graph_mode;                          // dpe_set_mcmc_graph: 0 eager launches, 1 replay a captured CUDA graph when a call repeats
    bool profile;
    struct ProfRec { cudaEvent_t e0, e1; int klass; double flops; int stage; };
    std::vector<ProfRec> *prof;
    int last_gemm_class;
    void *tc;      // dpe::TcState (gemm_tc.cu): tf32-split transposed weights + their TMA descriptors
    int n_sm;
};

namespace dpe {
#ifdef __CUDACC__
// log|det| = sum_p log|pivot_p| without a double-precision log per pivot (the FP64 pipe issues one warp instruction every other clock and
// log costs ~50 of them: 14 logs were 8x the elimination itself): the pivots are multiplied into a mantissa in [1, 2) with the binary exponent
// tracked separately, one log at the end.
struct LogDetAcc {
    double mant = 1.0;
    int ex = 0;
    __device__ __forceinline__ void mul(double piv) {
        mant *= fabs(piv);
        const long long bits = __double_as_longlong(mant);
        const int e = (int)((bits >> 52) & 0x7ff);
        if (e != 0 && e != 0x7ff) {                    // zero / denormal / inf / nan: leave as is, the final log reports it
            ex += e - 1023;
            mant = __longlong_as_double((bits & ~(0x7ffLL << 52)) | (1023LL << 52));
        }
    }
    __device__ __forceinline__ double value() const { return log(mant) + (double)ex * 0.69314718055994530942; }
};
#endif
void mcmc_graphs_destroy(dpe_model *m);

int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
#define DPE_CUDA(x) do { int _e = dpe::check_cuda((x), #x); if (_e) return _e; } while (0)
#define DPE_LAUNCH_CHECK(m) do { (m)->launches++; int _e = dpe::check_cuda(cudaGetLastError(), __func__); if (_e) return _e; } while (0)

// Kernels that need more than 48 KB of dynamic shared memory are opted in once per model (function attributes are per device /
// context, and a model lives on one device): always to the full 227 KB, so that two models with different needs cannot undercut
// each other.
enum KernelId { KID_PAIR1 = 0, KID_PAIR3, KID_CONV2_BIG, KID_CONV1, KID_DET64, KID_DET64F, KID_DET128, KID_GEMM_TC, KID_GEMM_TC2F,
                KID_GEMM_TC2P, KID_GEMM_ROWS, KID_DET_TRACE, KID_MCMC_FUSED, KID_GRAD_A, KID_GRAD_B, KID_PAIR_TC, KID_BW_PAIR, KID_BW_EION, KID_BW_PAIR_ROWS, KID_COUNT };
constexpr int DPE_SMEM_OPTIN = 227 * 1024;
template <typename F>
inline int opt_in_smem(dpe_model *m, int kid, F *fn) {
    if (m->smem_opted & (1ull << kid)) return DPE_OK;
    cudaFuncAttributes fa;
    int e = check_cuda(cudaFuncGetAttributes(&fa, fn), "cudaFuncGetAttributes");
    if (e) return e;
    // the opt-in limit covers static + dynamic shared memory
    e = check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, DPE_SMEM_OPTIN - (int)fa.sharedSizeBytes), "cudaFuncSetAttribute(smem)");
    if (e) return e;
    m->smem_opted |= 1ull << kid;
    return DPE_OK;
}

// Stage timing (bench.py / tools): while dpe_profile_enable is on, every launch group of run_chunk is bracketed by CUDA events on the
// launching stream.  Stage ids are the indices of DPE_STAGE_NAMES (include/dpe_b200.h).
enum Stage { ST_FEATURES = 0, ST_EION, ST_PAIR, ST_HMAP, ST_CONV, ST_MEAN, ST_MEAN_GEMM, ST_MAIN, ST_ORBITALS, ST_DET_FACTOR, ST_DET_TRACE,
             ST_COMBINE, ST_MCMC, ST_COUNT };
struct StageTimer {
    dpe_model *m; cudaStream_t s; dpe_model::ProfRec rec; bool on;
    StageTimer(dpe_model *m_, int stage, cudaStream_t s_) : m(m_), s(s_), on(m_->profile) {
        if (!on) return;
        rec.klass = -1; rec.flops = 0.0; rec.stage = stage;
        on = cudaEventCreate(&rec.e0) == cudaSuccess && cudaEventCreate(&rec.e1) == cudaSuccess && cudaEventRecord(rec.e0, s) == cudaSuccess;
    }
    ~StageTimer() {
        if (on && cudaEventRecord(rec.e1, s) == cudaSuccess) m->prof->push_back(rec);
    }
};

// gemm_simt.cu
int launch_gemm_simt(dpe_model *m, const GemmArgs &g, cudaStream_t s);
// gemm_tc.cu (tcgen05 3xTF32); returns DPE_ERR_UNSUPPORTED when the shape does not fit, caller falls back to SIMT
int launch_gemm_tc(dpe_model *m, const GemmArgs &g, cudaStream_t s);
int tc_register_weight(dpe_model *m, const float *W, int K, int N, bool tr = false);   // W: [K, N] row-major, device (tr: stored [N, K], used as W^T)
int tc_refresh_weights(dpe_model *m, cudaStream_t s);                  // re-split after a parameter update
void tc_destroy(dpe_model *m);

// streams.cu
int launch_features(dpe_model *m, const float *r, int Bc, int C, float *x0, int ldx, float *epot, cudaStream_t s);
int launch_eion_stream(dpe_model *m, const float *r, int Bc, int CE, float *ei_base, const size_t *ei_off_floats, cudaStream_t s);
int launch_pair_stream(dpe_model *m, const float *r, int Bc, int CP, float *pw_base, const size_t *pw_off_floats, cudaStream_t s);
int launch_act(dpe_model *m, float *z, int ld, int n_groups, int C, int width, const float *bias, const float *add,
               int groups_per_add, cudaStream_t s);
int launch_mean(dpe_model *m, const float *x, int ldx, int Bc, int C, int d_in, float *mean, cudaStream_t s);
int launch_act_mean_fwd(dpe_model *m, float *z, int ld, int Bc, int width, const float *bias, const float *add, float *mean, cudaStream_t s);
int launch_conv(dpe_model *m, int it, const float *r, int Bc, int C, const float *hm, const float *pw, const float *ei,
                float *x, int ldx, cudaStream_t s);
int launch_prepare_params(dpe_model *m, cudaStream_t s);
int launch_prepare_geometry(dpe_model *m, cudaStream_t s);
int launch_geometry_from_device(dpe_model *m, const float *R_dev, const int32_t *Z_dev, cudaStream_t s);

// orbitals_det.cu
int launch_envelope(dpe_model *m, const float *r, int Bc, int C, float *mo, cudaStream_t s);
// TAO orbitals (transferable_atomic_orbitals.py:287-349): mo = sum_ion g[ion] * exp(-exponent * |r_i - R_ion|) with the product rule
int launch_tao_pack(dpe_model *m, const float *bf_up, const float *bf_dn, const float *ex_up, const float *ex_dn, cudaStream_t s);
int launch_tao_orbitals(dpe_model *m, const float *r, int Bc, int C, const float *g, float *mo, cudaStream_t s);
int launch_det(dpe_model *m, int Bc, int C, const float *mo, float *det, float *ainv, cudaStream_t s);
// CTA-pair (cta_group::2) dense-layer kernel: 0 = off, 1 = plain launches only, 2 (default) = also the fused bias + tanh-rule epilogue
inline int tc_pair_mode() {
    static const int mode = getenv("DPE_TC_2CTA") ? atoi(getenv("DPE_TC_2CTA")) : 2;
    return mode;
}
// padded size of the Ainv^T tiles of the tensor-core determinant stage: room for the (det * N) mod 4 column shift that
// keeps the TMA box start 16-byte aligned, rounded up to the 16-float K slab
inline int det_tc_pad(int N, int n_det) {
    int omax = 0;
    for (int dt = 0; dt < n_det && dt < 4; ++dt) omax = ((dt * N) & 3) > omax ? ((dt * N) & 3) : omax;
    return (N + omax + 15) & ~15;
}
// det_tc.cu: traces of (dA_k Ainv) and (dA_k Ainv)^2 on the tensor cores; DPE_ERR_UNSUPPORTED if the shape does not fit
int launch_det_trace_tc(dpe_model *m, int Bc, int C, const float *mo, const float *ainv_hi, const float *ainv_lo, int NP, float *det, cudaStream_t s);
int launch_combine(dpe_model *m, int Bc, int C, const float *det, const float *epot, float *phase, float *logpsi2,
                   float *grad, float *ekin, float *eloc, float *epot_out, cudaStream_t s);

// api.cu: dense layers for grad.cu
int dense_gemm(dpe_model *m, const float *A, int lda, const float *W, float *Cc, int ldc, int M, int N, int K, cudaStream_t s);
// C = A W^T for a weight registered with tr (tensor cores only: DPE_ERR_UNSUPPORTED otherwise, the caller keeps its own kernel)
int dense_gemm_t(dpe_model *m, const float *A, int lda, const float *W, float *Cc, int ldc, int M, int N, int K, int seg_len, int seg_stride, int seg_off,
                 cudaStream_t s);
int dense_gemm_seg(dpe_model *m, const float *A, int lda, const float *W, float *Cc, int ldc, int M, int N, int K, int seg_len, int seg_stride, int seg_off,
                   cudaStream_t s);

// mcmc.cu
int launch_pair_stream_tc(dpe_model *m, const float *r, int Bc, float *pw_base, const size_t *pw_off, cudaStream_t s);
float tc_rz_comp(int K);
// gemm_tc.cu: split-K  At Bt^T  products of the gradient / KFAC pass on the CTA-pair tensor-core kernel (grad.cu prepares the transposed operands)
int launch_atb_tc(dpe_model *m, const float *At, float *Bt_hi, float *Bt_lo, long Rp, int Kc, int Mt, int Nb, float *part, cudaStream_t s);
int launch_propose(const dpe_model *m, const dpe_mcmc_state *st, int B, const dpe_mcmc_config &cfg, int step_offset, float *r_prop, float *thr, uint32_t *new_keys,
                   float *log_q, cudaStream_t s);
int launch_accept(const dpe_mcmc_state *st, int B, int n_el, const float *r_prop, const float *lp_prop, const float *thr,
                  const uint32_t *new_keys, const float *log_q, int max_age, int32_t *mask, int32_t *count, cudaStream_t s);
int launch_controller(const dpe_mcmc_state *st, const int32_t *counts, int n_steps, int64_t n_total,
                      const dpe_mcmc_config &cfg, cudaStream_t s);

}  // namespace dpe
