// XLA GPU custom-call entry points (legacy ABI of jaxlib 0.4.x, the reference's pinned jax 0.4.23):
//     void fn(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len)
// `buffers` = the operands followed by the results of the HLO custom_call, all device pointers; `opaque` = a descriptor serialised
// by the Python binding (include/dpe_b200.h: dpe_xla_descriptor).  These are thin: every one forwards to the C-ABI entry point
// the ctypes binding uses, so the jax-side stub in INTEGRATION.md B needs nothing but xla_client.register_custom_call_target.
// The legacy ABI cannot return a status: a failure is latched (dpe_xla_last_status / dpe_last_error) and the results are filled
// with NaN so that it cannot pass unnoticed (the reference stops on NaN, optimization.stop_on_nan).
#include <cstring>
#include "dpe_internal.cuh"

namespace dpe {

static int g_xla_status = DPE_OK;

__global__ void k_fill_nan(float *p, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = __int_as_float(0x7fc00000);
}

static void fail(int rc, float *out, long n, cudaStream_t s) {
    g_xla_status = rc;
    if (out && n > 0) k_fill_nan<<<(int)((n + 255) / 256), 256, 0, s>>>(out, n);
}

static const dpe_xla_descriptor *descriptor(const char *opaque, size_t len) {
    if (!opaque || len != sizeof(dpe_xla_descriptor)) {
        g_xla_status = set_error(DPE_ERR_ARG, "xla custom call: opaque has %zu bytes, expected sizeof(dpe_xla_descriptor) = %zu", len, sizeof(dpe_xla_descriptor));
        return nullptr;
    }
    return reinterpret_cast<const dpe_xla_descriptor *>(opaque);
}

}  // namespace dpe

using namespace dpe;

extern "C" {

int dpe_xla_last_status(void) { int s = g_xla_status; g_xla_status = DPE_OK; return s; }

// operands: r[B, n_el, 3], workspace[u8];  results: phase[B], log_psi_sqr[B]
void dpe_xla_log_psi_sqr(void *stream, void **buffers, const char *opaque, size_t opaque_len) {
    const dpe_xla_descriptor *d = descriptor(opaque, opaque_len);
    if (!d) return;
    cudaStream_t s = (cudaStream_t)stream;
    dpe_model *m = reinterpret_cast<dpe_model *>(d->model);
    int rc = dpe_log_psi_sqr(m, (const float *)buffers[0], d->n_walkers, (float *)buffers[2], (float *)buffers[3], buffers[1], d->workspace_bytes, stream);
    if (rc) { fail(rc, (float *)buffers[3], d->n_walkers, s); fail(rc, (float *)buffers[2], d->n_walkers, s); }
}

// operands: r[B, n_el, 3], workspace[u8];  results: e_loc[B]
void dpe_xla_local_energy(void *stream, void **buffers, const char *opaque, size_t opaque_len) {
    const dpe_xla_descriptor *d = descriptor(opaque, opaque_len);
    if (!d) return;
    dpe_model *m = reinterpret_cast<dpe_model *>(d->model);
    int rc = dpe_local_energy(m, (const float *)buffers[0], d->n_walkers, (float *)buffers[2], nullptr, nullptr, nullptr, nullptr, buffers[1],
                              d->workspace_bytes, stream);
    if (rc) fail(rc, (float *)buffers[2], d->n_walkers, (cudaStream_t)stream);
}

// operands: r, log_psi_sqr, walker_age, rng_state, stepsize, step_nr, acc_rate (the batch / scalar fields of MCMCState, mcmc.py:20-33),
//           workspace[u8];
// results:  the same seven fields after d->n_steps Metropolis steps, accept_counts[n_steps] (int32).
// XLA hands out distinct result buffers: the state is copied operand -> result first and advanced in place there.
void dpe_xla_mcmc_steps(void *stream, void **buffers, const char *opaque, size_t opaque_len) {
    const dpe_xla_descriptor *d = descriptor(opaque, opaque_len);
    if (!d) return;
    cudaStream_t s = (cudaStream_t)stream;
    dpe_model *m = reinterpret_cast<dpe_model *>(d->model);
    const size_t B = d->n_walkers, N = m->dims.n_el;
    const size_t bytes[7] = {B * N * 3 * sizeof(float), B * sizeof(float), B * sizeof(int32_t), B * 2 * sizeof(uint32_t), sizeof(float), sizeof(int32_t), sizeof(float)};
    void **in = buffers, **out = buffers + 8;
    for (int k = 0; k < 7; ++k)
        if (out[k] != in[k] && cudaMemcpyAsync(out[k], in[k], bytes[k], cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
            fail(set_error(DPE_ERR_CUDA, "xla mcmc_steps: state copy failed"), (float *)out[0], (long)(B * N * 3), s);
            return;
        }
    dpe_mcmc_state st = {(float *)out[0], (float *)out[1], (int32_t *)out[2], (uint32_t *)out[3], (float *)out[4], (int32_t *)out[5], (float *)out[6]};
    int rc = dpe_mcmc_steps(m, &st, d->n_walkers, d->n_steps, &d->mcmc, d->recompute_log_psi, d->run_controller, (int32_t *)out[7], in[7],
                            d->workspace_bytes, stream);
    if (rc) fail(rc, (float *)out[0], (long)(B * N * 3), s);
}

}  // extern "C"
