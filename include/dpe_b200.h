/*
 * dpe_b200.h -- C ABI of the B200-native DeepErwin VMC inner loop (libdpe_b200.so).
 *
 * The reference (mdsunivie/deeperwin) has no FFI layer: the boundary it exposes is a set of Python
 * callables.  Each entry point below names the reference callable it replaces
 * (paths relative to /root/reference/src/deeperwin/).  deeperwin_b200/ re-creates those callables on
 * top of this ABI with ctypes; INTEGRATION.md shows the jax-side binding.
 *
 * Conventions: plain pointers and sizes only; every `*_dev` pointer is a CUDA device pointer,
 * everything else is host memory; all floating point is float32 (computation.float_precision,
 * configuration.py:1893), integers int32 / uint32.  Every call is asynchronous on `stream` (a
 * cudaStream_t passed as void*), returns 0 on success or a negative dpe_status, never throws and never
 * allocates device memory behind the caller's back: scratch comes from the caller-provided workspace
 * (size it with dpe_workspace_bytes).  A handle may be used by one host thread at a time.
 * There is no CPU fallback: without a CUDA device every compute entry point returns DPE_ERR_CUDA.
 */
#ifndef DPE_B200_H
#define DPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPE_MAX_ITER 8

typedef enum dpe_status {
    DPE_OK = 0,
    DPE_ERR_ARG = -1,         /* null pointer / inconsistent sizes */
    DPE_ERR_UNSUPPORTED = -2, /* model shape outside what the kernels implement */
    DPE_ERR_WORKSPACE = -3,   /* workspace too small for even one walker */
    DPE_ERR_CUDA = -4,        /* a CUDA runtime call failed; see dpe_last_error() */
    DPE_ERR_STATE = -5        /* parameters / geometry not set yet */
} dpe_status;

/* Shape of the default `dpe4` model: ModelConfigDeepErwin4, configuration.py:422-440, 634-674,
 * 783-806, 875-890 (full_det determinants, isotropic exponential envelopes, tanh MLPs). */
typedef struct dpe_dims {
    int32_t n_el, n_up, n_ion;
    int32_t n_iterations;              /* embedding.n_iterations (4) */
    int32_t n_hidden_one_el[DPE_MAX_ITER]; /* [n_iterations]  (256) */
    int32_t n_hidden_two_el[DPE_MAX_ITER]; /* [n_iterations-1] (32); also used by the el-ion stream, ferminet_embedding.py:260 */
    int32_t emb_dim;                   /* 32 */
    int32_t n_ion_features;            /* 32 */
    int32_t n_dets;                    /* 32 */
    int32_t z_min, z_max;              /* lookup-embedding vocabulary, model/wavefunction.py:310-326 */
    int32_t use_taos;                  /* 0: envelope orbitals (orbitals/envelope_orbitals.py:39-127, dpe4 default);
                                          1: transferable atomic orbitals from a per-geometry cache
                                             (orbitals/transferable_atomic_orbitals.py:287-349, orbitals.envelope_orbitals = null) */
} dpe_dims;

/* MCMCConfig fields the Metropolis step reads (configuration.py:978-1049), `normal` proposal. */
typedef struct dpe_mcmc_config {
    int32_t max_age;
    int32_t stepsize_update_interval;
    float target_acceptance_rate;
    float min_stepsize_scale;
    float max_stepsize_scale;
    int32_t proposal;        /* proposal.name (configuration.py:952-975, mcmc.py:330-343): 0 "normal" (mcmc.py:175-180), 1 "cauchy" (:196-201),
                                2 "normal_one_el" (:183-193: electron step_nr % n_el moves) -- log_q_ratio = 0;
                                3 "local" (:212-228), 4 "local_one_el" (:231-253): step size x clip(distance to the closest nucleus, r_min, r_max);
                                5 "langevin" (:205-209, 256-284): local step size + drift -langevin_scale sum_J Z_J (r - R_J) / |r - R_J|;
                                3-5 carry their log_q_ratio into the acceptance probability (mcmc.py:359) */
    float r_min;             /* proposals 3-5 (LocalStepsizeProposalConfig / MCMCLangevinProposalConfig) */
    float r_max;
    float langevin_scale;    /* proposal 5 */
} dpe_mcmc_config;

/* Device-resident walker state: the batch-axis fields of MCMCState (mcmc.py:20-33, 149-151). */
typedef struct dpe_mcmc_state {
    float *r_dev;            /* [B, n_el, 3] */
    float *log_psi_sqr_dev;  /* [B] */
    int32_t *walker_age_dev; /* [B] */
    uint32_t *rng_state_dev; /* [B, 2] threefry keys */
    float *stepsize_dev;     /* [1] */
    int32_t *step_nr_dev;    /* [1] */
    float *acc_rate_dev;     /* [1] */
} dpe_mcmc_state;

typedef struct dpe_model dpe_model; /* opaque */

/* ---- lifetime -------------------------------------------------------------------------------- */
const char *dpe_version(void);
const char *dpe_last_error(void);

/* Replaces hk.multi_transform(Wavefunction(...)) of build_log_psi_squared (model/wavefunction.py:262-299):
 * fixes the shapes; weights arrive through dpe_model_set_params. */
int dpe_model_create(const dpe_dims *dims, dpe_model **out);
void dpe_model_destroy(dpe_model *m);

/* Number of float32 values in the flat parameter vector, and the offset/size of one leaf.
 * Canonical leaf order (the haiku tree of SURVEY.md 8b, flattened):
 *   0: wf/~/input/h_ion embeddings [V, n_ion_features]
 *   per iteration it: w_same.{w,b}, w_diff.{w,b}, h_map.{w,b}, h_ion_map.{w,b}, h_el_it.{w,b},
 *                     and for it < n_iterations-1: h_same_it.{w,b}, h_diff_it.{w,b}, h_el_ion_it.{w,b}
 *   then bf_up.w, bf_dn.w, alpha_up, alpha_dn, weights_up, weights_dn.
 * All `w` are [d_in, d_out] row-major (y = x @ w + b, hk.Linear). */
int64_t dpe_param_count(const dpe_model *m);
int32_t dpe_param_leaf_count(const dpe_model *m);
int dpe_param_leaf(const dpe_model *m, int32_t leaf, int64_t *offset, int64_t *size, int32_t *rows, int32_t *cols);

/* `params` of log_psi_sqr(params, ...) (model/wavefunction.py:293): copies the flat vector (device
 * memory) and rebuilds the derived kernel-side layouts. */
int dpe_model_set_params(dpe_model *m, const float *params_dev, int64_t n, void *stream);

/* `R, Z` of log_psi_sqr(params, n_up, n_dn, r, R, Z, fixed_params): host arrays R[n_ion*3], Z[n_ion].
 * The arrays are staged through pinned memory owned by the handle: the call does not synchronise the stream. */
int dpe_model_set_geometry(dpe_model *m, const float *R_host, const int32_t *Z_host, void *stream);

/* The same with R / Z already on the device (MCMCState.R float32[n_ion,3], MCMCState.Z int32[n_ion], mcmc.py:23-24): no host
 * round trip at all -- the weight-sharing loop switches geometry every optimisation step (variational_optimization.py:356-387).
 * Z outside [z_min, z_max] cannot be reported synchronously: it is clamped and latched; dpe_model_geometry_status (which
 * synchronises) returns DPE_ERR_ARG if that ever happened. */
int dpe_model_set_geometry_dev(dpe_model *m, const float *R_dev, const int32_t *Z_dev, void *stream);
int dpe_model_geometry_status(dpe_model *m, void *stream);

/* `fixed_params["cache"]["taos"]` of log_psi_sqr(...) for models built with use_taos = 1 (orbital_net.py:84-95,
 * model/wavefunction.py:164-209): the geometry-only outputs of TAOBackflow / TAOExponents, device arrays,
 *   backflows_{up,dn}  [n_ion, n_orb, 2, n_dets, n_hidden_one_el[last]]   (n_orb = n_up / n_dn orbitals of that spin)
 *   exponents_{up,dn}  [n_ion, n_orb, 2, n_dets]
 * axis 2 = (same spin, different spin) of electron vs orbital.  As in the reference with use_el_ion_embedding = False,
 * the backflow of BOTH spin types is slice 0 (transferable_atomic_orbitals.py:255-260); the exponents use both slices.
 * With use_taos = 1 the flat parameter vector ends after the embedding leaves (no bf_ / alpha_ / weights_ leaves). */
int dpe_model_set_tao_cache(dpe_model *m, const float *backflows_up_dev, const float *backflows_dn_dev,
                            const float *exponents_up_dev, const float *exponents_dn_dev, void *stream);

/* ---- workspace ------------------------------------------------------------------------------- */
/* mode 0: forward only (log psi^2), mode 1: forward-Laplacian (E_loc). Bytes needed to process
 * `n_walkers` walkers in one pass; calls given less process the batch in chunks. */
#define DPE_MODE_FORWARD 0
#define DPE_MODE_LAPLACIAN 1
size_t dpe_workspace_bytes(const dpe_model *m, int32_t n_walkers, int32_t mode);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Replaces log_psi_sqr (model/wavefunction.py:118-134, 293): r[B,n_el,3] -> phase[B] (0 or pi),
 * log_psi_sqr[B]. phase_dev may be NULL. */
int dpe_log_psi_sqr(dpe_model *m, const float *r_dev, int32_t n_walkers, float *phase_dev, float *log_psi_sqr_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream);

/* Replaces get_local_energy built by build_local_energy(..., forward_lap=True) (hamiltonian.py:272-291,
 * kinetic part :206-216, potential :34-39). Optional outputs (NULL to skip): log_psi_sqr[B],
 * grad_log_psi_sqr[B,3*n_el], E_kin[B], E_pot[B]. */
int dpe_local_energy(dpe_model *m, const float *r_dev, int32_t n_walkers, float *e_loc_dev, float *log_psi_sqr_dev,
                     float *grad_dev, float *e_kin_dev, float *e_pot_dev, void *workspace_dev, size_t workspace_bytes,
                     void *stream);

/* Replaces MetropolisHastingsMonteCarlo._run_mcmc_steps (mcmc.py:389-406) with the `normal` proposal
 * (mcmc.py:175-180) and make_mcmc_step (mcmc.py:345-379) for one device's walkers.
 *  - recompute_log_psi != 0 re-evaluates state.log_psi_sqr first (mcmc.py:396);
 *  - accept_counts_dev[n_steps] (int32) receives the number of accepted walkers of each step;
 *  - run_controller != 0 applies the acceptance-rate EMA / step-size controller (mcmc.py:367-377) after
 *    every step using this device's counts over n_walkers_total == n_walkers (single GPU). With several
 *    GPUs pass 0, all-reduce accept_counts_dev and call dpe_mcmc_controller; n_steps must then not cross
 *    a multiple of stepsize_update_interval (the host shim segments the call). */
int dpe_mcmc_steps(dpe_model *m, const dpe_mcmc_state *state, int32_t n_walkers, int32_t n_steps,
                   const dpe_mcmc_config *cfg, int32_t recompute_log_psi, int32_t run_controller,
                   int32_t *accept_counts_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* The scalar tail of make_mcmc_step (mcmc.py:367-377) replayed for n_steps steps from (all-reduced)
 * accept counts: acc_rate EMA, step_nr += 1, step-size update with the pre-update acc_rate. */
int dpe_mcmc_controller(const dpe_mcmc_state *state, const int32_t *accept_counts_dev, int32_t n_steps,
                        int64_t n_walkers_total, const dpe_mcmc_config *cfg, void *stream);

/* Replaces the statistics of total_energy (optimization/loss_function.py:89-109) on one device.
 * Pass 1 (dpe_energy_moments1): out[0]=nanmean(E), out[1]=nanmean(E_clipped), writes E_clipped[B]
 *   using clip center/width (tanh: clip_mode 0, hard: 1).
 * Pass 2 (dpe_energy_moments2): out[0]=nanmean((E-e_mean)^2), out[1]=nanmean((E_clipped-c_mean)^2),
 *   reading e_mean / c_mean from means_dev[0..1] (after the caller's all-reduce). */
int dpe_energy_moments1(const float *e_loc_dev, int32_t n, const float *clip_center_width_dev, int32_t clip_mode,
                        float *e_clipped_dev, float *out2_dev, void *stream);
int dpe_energy_moments2(const float *e_loc_dev, const float *e_clipped_dev, int32_t n, const float *means_dev,
                        float *out2_dev, void *stream);
/* The other clipping-window statistics of _get_clipping_center_and_width (loss_function.py:19-30), per device:
 * dpe_energy_median: out[0] = nanmedian(E) (clipping.center = "median"; the caller all-reduces the per-device medians
 *   exactly as the reference's pmean(nanmedian) does);
 * dpe_energy_width: out[0] = nanmean((E - center)^2) (metric 0, "std": the caller all-reduces, then sqrt) or
 *   nanmean(|E - center|) (metric 1, "mae"), center read from center_dev[0]. */
int dpe_energy_median(const float *e_dev, int32_t n, float *out_dev, void *stream);
int dpe_energy_width(const float *e_dev, int32_t n, const float *center_dev, int32_t metric, float *out_dev, void *stream);

/* Repeated dpe_mcmc_steps calls (identical pointers, sizes, config) are replayed from a CUDA graph captured on the second occurrence; mode 0 keeps
 * every call on plain launches (also: environment DPE_MCMC_GRAPH=0).  Results are identical either way (same kernels, same order). */
int dpe_set_mcmc_graph(dpe_model *m, int32_t mode);
int dpe_get_mcmc_graph(const dpe_model *m);

/* ---- optimisation step: parameter gradient and KFAC statistics --------------------------------
 * dpe_param_gradient is the backward pass of log psi^2 on the value channel.
 *   grad_dev[n_params] (canonical leaf order, may be NULL) = sum_b cotangent_dev[b] * d log_psi_sqr_b / d params: with
 *     cotangent_b = (E_clipped_b - mean E_clipped) / B it is the gradient of total_energy (optimization/loss_function.py:143-154);
 *   kfac_dev[dpe_kfac_floats] (may be NULL) receives, per dense layer, the Kronecker factors of kfac_jax's dense blocks with the
 *     repeated-dense folding (custom_kfac_jax/kfac_jax/_src/curvature_blocks.py:1594-1624, curvature_tags_and_blocks.py:41-64) for the
 *     loss registered in loss_function.py:148-150 (normal predictive distribution on 1/2 log psi^2, variance 1/2, fisher_exact):
 *        A[(din + bias)^2] = [x, 1]^T [x, 1] / B',  G[dout^2] = dy^T dy / B',  dy = (1 / sqrt 2) d log psi^2 / dy,  B' = n_walkers * rows_per_walker.
 *     The layers, their shapes and offsets: dpe_kfac_layer_count / dpe_kfac_layer (name = the haiku module of the layer);
 *   log_psi_sqr_dev[B] (may be NULL) receives log psi^2 of the same pass.
 * TAO models (use_taos): the gradient covers the embedding leaves (the flat vector has no orbital leaves), the factor list has no orbital layers, and
 * the geometry cache of dpe_model_set_tao_cache is held fixed (its cotangents belong to the geometry-only nets, outside this library).
 * Both outputs are sums / means over THIS device's walkers: with several GPUs the caller all-reduces the two buffers (one flat
 * all-reduce, optimizers.py:133, kfac optimizer.py:1151).  The workspace (dpe_gradient_workspace_bytes) holds the saved activations of a
 * chunk of walkers, the split-K partial products and the transposed operands of the tensor-core products (N2 x 4096 walkers: 3.4 GB);
 * batches larger than the workspace handed in are processed in chunks that accumulate. */
int32_t dpe_kfac_layer_count(const dpe_model *m);
int64_t dpe_kfac_floats(const dpe_model *m);
int dpe_kfac_layer(const dpe_model *m, int32_t index, char *name, int32_t name_len, int32_t *din, int32_t *dout, int32_t *has_bias,
                   int32_t *rows_per_walker, int64_t *a_offset, int64_t *g_offset);
size_t dpe_gradient_workspace_bytes(const dpe_model *m, int32_t n_walkers);
int dpe_param_gradient(dpe_model *m, const float *r_dev, int32_t n_walkers, const float *cotangent_dev, float *grad_dev, float *kfac_dev,
                       float *log_psi_sqr_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- XLA custom-call entry points (keeping JAX as the host) ------------------------------------------
 * Legacy GPU custom-call ABI of the reference's pinned jaxlib (jax 0.4.23): void fn(cudaStream_t, void **buffers, const char
 * *opaque, size_t opaque_len) with `buffers` = operands then results (device pointers).  They replace the three switch points
 * of the reference: log_psi_sqr (model/wavefunction.py:293), get_local_energy (hamiltonian.py:280-289) and the pmapped
 * _run_mcmc_steps (mcmc.py:325-327, 389-406).  `opaque` is one dpe_xla_descriptor, serialised by the Python binding.
 *   dpe_xla_log_psi_sqr : operands r, workspace                      results phase, log_psi_sqr
 *   dpe_xla_local_energy: operands r, workspace                      results e_loc
 *   dpe_xla_mcmc_steps  : operands r, log_psi_sqr, walker_age, rng_state, stepsize, step_nr, acc_rate, workspace
 *                         results  the same seven MCMCState fields after n_steps, accept_counts[n_steps]
 * The legacy ABI has no return value: a failure fills the results with NaN and is latched; dpe_xla_last_status() returns and
 * clears it (0 if every custom call since the last query succeeded), dpe_last_error() holds the message. */
typedef struct dpe_xla_descriptor {
    uint64_t model;            /* dpe_model* of dpe_model_create, as an integer */
    uint64_t workspace_bytes;  /* size of the workspace operand */
    int32_t n_walkers;
    int32_t n_steps;           /* mcmc_steps only */
    int32_t recompute_log_psi; /* mcmc_steps only: mcmc.py:396 */
    int32_t run_controller;    /* mcmc_steps only: 1 on a single device, 0 when the caller psum-s accept_counts first */
    dpe_mcmc_config mcmc;      /* mcmc_steps only */
} dpe_xla_descriptor;
void dpe_xla_log_psi_sqr(void *stream, void **buffers, const char *opaque, size_t opaque_len);
void dpe_xla_local_energy(void *stream, void **buffers, const char *opaque, size_t opaque_len);
void dpe_xla_mcmc_steps(void *stream, void **buffers, const char *opaque, size_t opaque_len);
int dpe_xla_last_status(void);

/* ---- test hooks (jax.random restated; oracle/threefry.py) ------------------------------------ */
/* keys[B,2] -> new_keys[B,2], noise[B,n,3]=normal(sub,[n,3]), thr[B]=uniform(sub,()), as one Metropolis
 * step consumes them (mcmc.py:178-179, 360-361). */
int dpe_threefry_mcmc_randoms(const uint32_t *keys_dev, int32_t n_walkers, int32_t n_el, uint32_t *new_keys_dev,
                              float *noise_dev, float *thr_dev, void *stream);
/* bits(key, n): jax _threefry_random_bits for one key (host key[2]) -> bits_dev[n]. */
int dpe_threefry_bits(const uint32_t *key_host, int32_t n, uint32_t *bits_dev, void *stream);

/* normal(key, [n]) for one key (host key[2]) -> out_dev[n] float32 (walker initialisation, mcmc.py:65). */
int dpe_threefry_normal(const uint32_t *key_host, int32_t n, float *out_dev, void *stream);

/* Debug: byte offset of a named intermediate inside the workspace for a chunk of n_walkers in `mode`
 * (names: "x0".."x7","hm","mean","add","pw","ei","mo","det","epot"); -1 if unknown. */
int64_t dpe_debug_ws_offset(const dpe_model *m, int32_t n_walkers, int32_t mode, const char *name);
/* Which GEMM path the library was built to use for the dense layers: 0 = FP32 SIMT, 1 = tcgen05 3xTF32. */
int dpe_set_gemm_path(dpe_model *m, int32_t path);
/* Test hook: C[row(m), c_col_off + n] = sum_k A[row(m), k] W[k, n] through GEMM path 0 / 1 with the library's segmented
 * row addressing (row m -> (m / seg_len) * seg_stride + seg_off + m % seg_len). W is [K, N] row-major. */
int dpe_debug_gemm(dpe_model *m, int32_t path, const float *a_dev, int32_t lda, const float *w_dev, float *c_dev, int32_t ldc,
                   int32_t M, int32_t N, int32_t K, int32_t seg_len, int32_t a_seg_stride, int32_t a_seg_off, int32_t c_seg_stride,
                   int32_t c_seg_off, int32_t c_col_off, void *stream);
int dpe_get_gemm_path(const dpe_model *m);
/* Which kernels the determinant stage (model/wavefunction.py:63-83) uses -- test knob, every combination computes the same
 * quantities: bit 0 = the generic block-per-matrix kernel also for n_el <= 16 (default there: one warp per matrix),
 * bit 1 = tangent / Laplacian traces on CUDA cores instead of the tcgen05 trace kernel. Default 0. */
int dpe_set_det_path(dpe_model *m, int32_t flags);
int dpe_get_det_path(const dpe_model *m);
/* Live kernel timing for bench.py's roofline: while enabled, every dense-layer GEMM launch is bracketed by
 * CUDA events on the launching stream. dpe_profile_collect synchronises and returns, for GEMM kernel class
 * `klass` (0: SIMT 128x128, 1: SIMT 128x64, 2: SIMT 256x32, 3: tcgen05 3xTF32), the summed device time (ms), the
 * number of launches and their algorithmic FLOPs (2*M*N*K each); it then clears the records. */
int dpe_profile_enable(dpe_model *m, int32_t on);
int dpe_profile_collect(dpe_model *m, int32_t klass, double *ms, int64_t *count, double *flops);
/* Same, but per launch: fills ms_arr / flops_arr (capacity cap) in launch order and sets *n; does not clear the records. */
int dpe_profile_launches(dpe_model *m, int32_t klass, double *ms_arr, double *flops_arr, int32_t cap, int32_t *n);
/* While profiling is enabled every launch group of a pass is also timed per STAGE; dpe_profile_stages synchronises, fills
 * ms_arr / count_arr (capacity cap >= the number of stages) with the summed device time and the number of timed groups per
 * stage, clears those records and returns the number of stages (negative dpe_status on error).  Stage order: */
#define DPE_STAGE_NAMES "features", "el_ion_stream", "pair_stream", "h_map", "schnet_conv", "spin_mean", "mean_term_gemm", "main_layer", \
                        "orbitals", "det_factor", "det_trace", "combine", "mcmc"
int dpe_profile_stages(dpe_model *m, double *ms_arr, int64_t *count_arr, int32_t cap);
/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t dpe_launch_count(const dpe_model *m);

#ifdef __cplusplus
}
#endif
#endif /* DPE_B200_H */
