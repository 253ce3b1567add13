// Metropolis-Hastings walker update with a jax-compatible threefry2x32 counter RNG.
// Reference: mcmc.py:175-180 (_propose_normal), :345-387 (make_mcmc_step, _adjust_stepsize),
// utils/utils.py:115 (batch_rng_split); RNG = jax 0.4.23 threefry (un-vendored; restated in
// oracle/threefry.py and pinned by its known-answer tests).  Also the E_loc statistics of
// optimization/loss_function.py:19-30, 62-72, 89-109.
#include "dpe_internal.cuh"

namespace dpe {

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int d) { return (x << d) | (x >> (32 - d)); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t &x0, uint32_t &x1) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    x0 += ks[0]; x1 += ks[1];
#pragma unroll
    for (int g = 0; g < 5; ++g) {
        if ((g & 1) == 0) {
            x0 += x1; x1 = rotl32(x1, 13); x1 ^= x0;
            x0 += x1; x1 = rotl32(x1, 15); x1 ^= x0;
            x0 += x1; x1 = rotl32(x1, 26); x1 ^= x0;
            x0 += x1; x1 = rotl32(x1, 6);  x1 ^= x0;
        } else {
            x0 += x1; x1 = rotl32(x1, 17); x1 ^= x0;
            x0 += x1; x1 = rotl32(x1, 29); x1 ^= x0;
            x0 += x1; x1 = rotl32(x1, 16); x1 ^= x0;
            x0 += x1; x1 = rotl32(x1, 24); x1 ^= x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
}

// jax.random.split(key, 2): bits(key, 4) with counters (0,2) and (1,3) -> new_key = (y0a, y0b), sub = (y1a, y1b)
__device__ __forceinline__ void split2(uint32_t k0, uint32_t k1, uint32_t (&nk)[2], uint32_t (&sub)[2]) {
    uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
    threefry2x32(k0, k1, a0, a1);
    threefry2x32(k0, k1, b0, b1);
    nk[0] = a0; nk[1] = b0;
    sub[0] = a1; sub[1] = b1;
}

__device__ __forceinline__ float bits_to_unit(uint32_t bits) { return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f; }

// XLA's f32 erf_inv (Giles), evaluated without fused contractions to stay close to the CPU restatement
__device__ __forceinline__ float erf_inv_f32(float x) {
    float w = -log1pf(__fmul_rn(-x, x));
    float p;
    if (w < 5.0f) {
        w = __fsub_rn(w, 2.5f);
        p = 2.81022636e-08f;
        p = __fadd_rn(3.43273939e-07f, __fmul_rn(p, w));
        p = __fadd_rn(-3.5233877e-06f, __fmul_rn(p, w));
        p = __fadd_rn(-4.39150654e-06f, __fmul_rn(p, w));
        p = __fadd_rn(0.00021858087f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00125372503f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00417768164f, __fmul_rn(p, w));
        p = __fadd_rn(0.246640727f, __fmul_rn(p, w));
        p = __fadd_rn(1.50140941f, __fmul_rn(p, w));
    } else {
        w = __fsub_rn(sqrtf(w), 3.0f);
        p = -0.000200214257f;
        p = __fadd_rn(0.000100950558f, __fmul_rn(p, w));
        p = __fadd_rn(0.00134934322f, __fmul_rn(p, w));
        p = __fadd_rn(-0.00367342844f, __fmul_rn(p, w));
        p = __fadd_rn(0.00573950773f, __fmul_rn(p, w));
        p = __fadd_rn(-0.0076224613f, __fmul_rn(p, w));
        p = __fadd_rn(0.00943887047f, __fmul_rn(p, w));
        p = __fadd_rn(1.00167406f, __fmul_rn(p, w));
        p = __fadd_rn(2.83297682f, __fmul_rn(p, w));
    }
    return __fmul_rn(p, x);
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
    const float lo = -0.99999994f;                       // nextafter(-1, 0)
    float u = __fadd_rn(__fmul_rn(bits_to_unit(bits), 2.0f), lo);   // (hi - lo) rounds to 2.0f
    u = fmaxf(lo, u);
    return __fmul_rn(1.41421356237309515f, erf_inv_f32(u));
}

// jax.random.cauchy (jax 0.4.23): tan(pi * (uniform(minval = eps, maxval = 1) - 0.5)), all float32
__device__ __forceinline__ float bits_to_cauchy(uint32_t bits) {
    const float eps = 1.1920929e-07f;
    float u = __fadd_rn(__fmul_rn(bits_to_unit(bits), __fsub_rn(1.0f, eps)), eps);
    u = fmaxf(eps, u);
    return tanf(__fmul_rn(3.14159274f, __fsub_rn(u, 0.5f)));
}

// "normal_one_el" (mcmc.py:183-193): only electron step_nr % n_el moves, by normal(sub, [3]) * stepsize.  jax lays the three
// values out as bits(sub, 3): counters (0, 2) -> elements 0 and 2, counters (1, pad 0) -> element 1.  One thread per coordinate.
__global__ void __launch_bounds__(256) k_propose_one_el(const float *__restrict__ r, const uint32_t *__restrict__ keys,
                                                         const float *__restrict__ stepsize, const int32_t *__restrict__ step_nr,
                                                         int step_offset, int B, int n_el, float *__restrict__ r_prop, float *__restrict__ thr,
                                                         uint32_t *__restrict__ new_keys) {
    const int n = 3 * n_el;
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= (long)B * n) return;
    const long b = idx / n;
    const int e = (int)(idx - b * n), el = e / 3, k = e - 3 * el;
    const int moved = (step_nr[0] + step_offset) % n_el;
    float v = r[idx];
    if (el == moved || e == 0) {
        uint32_t nk[2], sub[2];
        split2(keys[2 * b], keys[2 * b + 1], nk, sub);
        if (el == moved) {
            uint32_t x0 = (k == 1) ? 1u : 0u, x1 = (k == 1) ? 0u : 2u;
            threefry2x32(sub[0], sub[1], x0, x1);
            const float nz = bits_to_normal(k == 2 ? x1 : x0);
            v = stepsize ? __fadd_rn(v, __fmul_rn(nz, stepsize[0])) : nz;        // stepsize == nullptr: the raw noise (local_one_el)
        }
        if (e == 0) {
            uint32_t t0 = 0u, t1 = 0u;
            threefry2x32(sub[0], sub[1], t0, t1);
            thr[b] = fmaxf(0.f, bits_to_unit(t0));
            new_keys[2 * b] = nk[0];
            new_keys[2 * b + 1] = nk[1];
        }
    }
    r_prop[idx] = v;
}

// One thread per (walker, counter pair p): noise element p from y0, element h+p from y1 (jax bits layout).
__global__ void __launch_bounds__(256) k_propose(const float *__restrict__ r, const uint32_t *__restrict__ keys,
                                                  const float *__restrict__ stepsize, int B, int n, float *__restrict__ r_prop,
                                                  float *__restrict__ noise_out, float *__restrict__ thr,
                                                  uint32_t *__restrict__ new_keys, int cauchy) {
    const int h = (n + 1) / 2;
    const long total = (long)B * h;
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long b = idx / h;
    const int p = (int)(idx - b * h);
    uint32_t nk[2], sub[2];
    split2(keys[2 * b], keys[2 * b + 1], nk, sub);
    uint32_t x0 = (uint32_t)p, x1 = (h + p < n) ? (uint32_t)(h + p) : 0u;   // odd n: padded counter is 0
    threefry2x32(sub[0], sub[1], x0, x1);
    const float ss = stepsize ? stepsize[0] : 1.f;
    float n0 = cauchy ? bits_to_cauchy(x0) : bits_to_normal(x0);
    if (noise_out) noise_out[b * n + p] = n0;
    if (r_prop) r_prop[b * n + p] = __fadd_rn(r[b * n + p], __fmul_rn(n0, ss));
    if (h + p < n) {
        float n1 = cauchy ? bits_to_cauchy(x1) : bits_to_normal(x1);
        if (noise_out) noise_out[b * n + h + p] = n1;
        if (r_prop) r_prop[b * n + h + p] = __fadd_rn(r[b * n + h + p], __fmul_rn(n1, ss));
    }
    if (p == 0) {
        uint32_t t0 = 0u, t1 = 0u;                    // uniform(sub, ()) = bits(sub, 1)[0]: counters (0, 0)
        threefry2x32(sub[0], sub[1], t0, t1);
        thr[b] = fmaxf(0.f, bits_to_unit(t0));
        new_keys[2 * b] = nk[0];
        new_keys[2 * b + 1] = nk[1];
    }
}

// ---- proposals with a position-dependent step size (mcmc.py:204-284) -----------------------------------------------------------
// s(r_i) = stepsize * clip(min_J |r_i - R_J|, r_min, r_max);  langevin adds the drift g(r_i) = -scale sum_J Z_J (r_i - R_J) / |r_i - R_J| times s^2.
// The arithmetic follows the reference expression by expression in float32 without FMA contraction (positions must not depend on it).
struct LocalStep { float s, g[3]; };
__device__ __forceinline__ LocalStep local_step(const float *ri, const float *__restrict__ R, const float *__restrict__ Z, int n_ion, float stepsize,
                                                float r_min, float r_max, float scale, bool langevin) {
    float dmin = 3.4e38f, gx = 0.f, gy = 0.f, gz = 0.f;
    for (int J = 0; J < n_ion; ++J) {
        const float dx = __fsub_rn(ri[0], R[3 * J]), dy = __fsub_rn(ri[1], R[3 * J + 1]), dz = __fsub_rn(ri[2], R[3 * J + 2]);
        const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        dmin = fminf(dmin, d);
        if (langevin) {                              // sum over ions of diff * Z / dist
            gx = __fadd_rn(gx, __fdiv_rn(__fmul_rn(dx, Z[J]), d));
            gy = __fadd_rn(gy, __fdiv_rn(__fmul_rn(dy, Z[J]), d));
            gz = __fadd_rn(gz, __fdiv_rn(__fmul_rn(dz, Z[J]), d));
        }
    }
    LocalStep o;
    o.s = __fmul_rn(stepsize, fminf(fmaxf(dmin, r_min), r_max));
    o.g[0] = __fmul_rn(-scale, gx); o.g[1] = __fmul_rn(-scale, gy); o.g[2] = __fmul_rn(-scale, gz);
    return o;
}

// log q(r | r') - log q(r' | r) of one electron: 3 (log s - log s') + (d_fwd / s^2 - d_rev / s'^2) / 2
__device__ __forceinline__ float log_q_term(float s, float s_new, float d_fwd, float d_rev) {
    const float a = __fmul_rn(3.f, __fsub_rn(logf(s), logf(s_new)));
    const float b = __fmul_rn(0.5f, __fsub_rn(__fdiv_rn(d_fwd, __fmul_rn(s, s)), __fdiv_rn(d_rev, __fmul_rn(s_new, s_new))));
    return __fadd_rn(a, b);
}

// "local" / "langevin": r_prop holds the raw noise of k_propose on entry, the proposed positions on exit.  One warp per walker (lanes over
// electrons), the per-electron log_q terms are summed in a fixed order.
__global__ void __launch_bounds__(256) k_local_finish(const float *__restrict__ r, const float *__restrict__ R, const float *__restrict__ Z, int n_ion,
                                                       const float *__restrict__ stepsize, int B, int n_el, float r_min, float r_max, float scale,
                                                       int langevin, float *__restrict__ r_prop, float *__restrict__ log_q) {
    const int lane = threadIdx.x & 31;
    const long b = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float ss = stepsize[0];
    float lq = 0.f;
    for (int i0 = 0; i0 < n_el; i0 += 32) {           // every lane walks the same number of rounds (shuffles below)
        const int i = i0 + lane;
        float term = 0.f;
        if (i < n_el) {
            const float *ri = r + (b * n_el + i) * 3;
            float *rp = r_prop + (b * n_el + i) * 3;
            const LocalStep o = local_step(ri, R, Z, n_ion, ss, r_min, r_max, scale, langevin);
            float rn[3], drift[3];
            for (int k = 0; k < 3; ++k) {
                drift[k] = langevin ? __fmul_rn(o.g[k], __fmul_rn(o.s, o.s)) : 0.f;
                rn[k] = __fadd_rn(ri[k], __fmul_rn(rp[k], o.s));
                if (langevin) rn[k] = __fadd_rn(rn[k], drift[k]);
            }
            const LocalStep n = local_step(rn, R, Z, n_ion, ss, r_min, r_max, scale, langevin);
            float d_fwd = 0.f, d_rev = 0.f;
            for (int k = 0; k < 3; ++k) {
                const float f = langevin ? __fsub_rn(__fsub_rn(rn[k], ri[k]), drift[k]) : __fsub_rn(rn[k], ri[k]);
                d_fwd = __fadd_rn(d_fwd, __fmul_rn(f, f));
                if (langevin) {
                    const float v = __fsub_rn(__fsub_rn(ri[k], rn[k]), __fmul_rn(n.g[k], __fmul_rn(n.s, n.s)));
                    d_rev = __fadd_rn(d_rev, __fmul_rn(v, v));
                }
                rp[k] = rn[k];
            }
            if (!langevin) d_rev = d_fwd;                 // local: 0.5 dist_sqr (1 / s^2 - 1 / s'^2)
            term = log_q_term(o.s, n.s, d_fwd, d_rev);
        }
        for (int off = 16; off; off >>= 1) term += __shfl_xor_sync(0xffffffffu, term, off);
        lq += term;
    }
    if (lane == 0) log_q[b] = lq;
}

// "local_one_el" (mcmc.py:231-253): k_propose_one_el has moved electron step_nr % n_el by noise * stepsize; rescale that move to the local
// step size of the electron and form its log_q_ratio.  One thread per walker.
__global__ void __launch_bounds__(256) k_local_one_el_finish(const float *__restrict__ r, const float *__restrict__ R, int n_ion,
                                                              const float *__restrict__ stepsize, const int32_t *__restrict__ step_nr, int step_offset,
                                                              int B, int n_el, float r_min, float r_max, float *__restrict__ r_prop,
                                                              float *__restrict__ log_q) {
    const long b = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int i = (step_nr[0] + step_offset) % n_el;
    const float *ri = r + (b * n_el + i) * 3;
    float *rp = r_prop + (b * n_el + i) * 3;                 // holds the raw noise (k_propose_one_el was run with stepsize = nullptr -> noise only)
    const float ss = stepsize[0];
    const LocalStep o = local_step(ri, R, nullptr, n_ion, ss, r_min, r_max, 0.f, false);
    float rn[3], d = 0.f;
    for (int k = 0; k < 3; ++k) rn[k] = __fadd_rn(ri[k], __fmul_rn(rp[k], o.s));
    const LocalStep n = local_step(rn, R, nullptr, n_ion, ss, r_min, r_max, 0.f, false);
    for (int k = 0; k < 3; ++k) {
        const float f = __fsub_rn(rn[k], ri[k]);
        d = __fadd_rn(d, __fmul_rn(f, f));
        rp[k] = rn[k];
    }
    log_q[b] = log_q_term(o.s, n.s, d, d);
}

int launch_propose(const dpe_model *m, const dpe_mcmc_state *st, int B, const dpe_mcmc_config &cfg, int step_offset, float *r_prop, float *thr,
                   uint32_t *new_keys, float *log_q, cudaStream_t s) {
    const int n_el = m->dims.n_el, n = 3 * n_el, h = (n + 1) / 2, proposal = cfg.proposal;
    if (proposal == 2 || proposal == 4) {
        long total = (long)B * n;
        // proposal 4: raw noise for the moved electron (unit step), positions copied for the others
        k_propose_one_el<<<(int)((total + 255) / 256), 256, 0, s>>>(st->r_dev, st->rng_state_dev, proposal == 2 ? st->stepsize_dev : nullptr, st->step_nr_dev,
                                                                    step_offset, B, n_el, r_prop, thr, new_keys);
        if (int e = check_cuda(cudaGetLastError(), "k_propose_one_el")) return e;
        if (proposal == 4) {
            k_local_one_el_finish<<<(B + 255) / 256, 256, 0, s>>>(st->r_dev, m->R_dev, m->dims.n_ion, st->stepsize_dev, st->step_nr_dev, step_offset, B, n_el,
                                                                  cfg.r_min, cfg.r_max, r_prop, log_q);
            return check_cuda(cudaGetLastError(), "k_local_one_el_finish");
        }
        return DPE_OK;
    }
    long total = (long)B * h;
    const bool local = proposal == 3 || proposal == 5;
    // local / langevin: k_propose leaves the raw noise in r_prop, k_local_finish turns it into positions and log_q_ratio
    k_propose<<<(int)((total + 255) / 256), 256, 0, s>>>(st->r_dev, st->rng_state_dev, st->stepsize_dev, B, n, local ? nullptr : r_prop, local ? r_prop : nullptr,
                                                         thr, new_keys, proposal == 1);
    if (int e = check_cuda(cudaGetLastError(), "k_propose")) return e;
    if (local) {
        k_local_finish<<<(int)(((long)B * 32 + 255) / 256), 256, 0, s>>>(st->r_dev, m->R_dev, m->Z_dev, m->dims.n_ion, st->stepsize_dev, B, n_el, cfg.r_min,
                                                                        cfg.r_max, cfg.langevin_scale, proposal == 5, r_prop, log_q);
        return check_cuda(cudaGetLastError(), "k_local_finish");
    }
    return DPE_OK;
}

// accept/reject (mcmc.py:358-366): one thread per walker decides, then the block copies positions.
__global__ void __launch_bounds__(256) k_accept(float *__restrict__ r, float *__restrict__ lp, int32_t *__restrict__ age,
                                                 uint32_t *__restrict__ keys, const float *__restrict__ r_prop,
                                                 const float *__restrict__ lp_prop, const float *__restrict__ thr,
                                                 const uint32_t *__restrict__ new_keys, const float *__restrict__ log_q, int B, int n, int max_age,
                                                 int32_t *__restrict__ mask, int32_t *__restrict__ count) {
    __shared__ int s_acc[256];
    __shared__ int s_cnt;
    const int b0 = blockIdx.x * blockDim.x;
    const int b = b0 + threadIdx.x;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int acc = 0;
    if (b < B) {
        float lo = lp[b], ln = lp_prop[b];
        float p_acc = log_q ? expf(__fadd_rn(__fsub_rn(ln, lo), log_q[b])) : expf(ln - lo);      // mcmc.py:359; log_q_ratio = 0 for proposals 0-2
        int a = age[b];
        acc = (p_acc > thr[b]) || (a >= max_age);
        age[b] = acc ? 0 : a + 1;
        if (acc) lp[b] = ln;
        keys[2 * b] = new_keys[2 * b];
        keys[2 * b + 1] = new_keys[2 * b + 1];
        if (mask) mask[b] = acc;
        if (acc) atomicAdd(&s_cnt, 1);
    }
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    const int nb = min((int)blockDim.x, B - b0);
    for (int e = threadIdx.x; e < nb * n; e += blockDim.x) {
        int w = e / n;
        if (s_acc[w]) r[(long)b0 * n + e] = r_prop[(long)b0 * n + e];
    }
    if (threadIdx.x == 0 && s_cnt) atomicAdd(count, s_cnt);
}

int launch_accept(const dpe_mcmc_state *st, int B, int n_el, const float *r_prop, const float *lp_prop, const float *thr,
                  const uint32_t *new_keys, const float *log_q, int max_age, int32_t *mask, int32_t *count, cudaStream_t s) {
    k_accept<<<(B + 255) / 256, 256, 0, s>>>(st->r_dev, st->log_psi_sqr_dev, st->walker_age_dev, st->rng_state_dev, r_prop,
                                             lp_prop, thr, new_keys, log_q, B, 3 * n_el, max_age, mask, count);
    return check_cuda(cudaGetLastError(), "k_accept");
}

// scalar tail of make_mcmc_step (mcmc.py:367-377), replayed over n_steps accept counts
__global__ void k_controller(float *stepsize, int32_t *step_nr, float *acc_rate, const int32_t *counts, int n_steps,
                             float total, dpe_mcmc_config cfg) {
    if (threadIdx.x || blockIdx.x) return;
    float ss = stepsize[0], ar = acc_rate[0];
    int sn = step_nr[0];
    for (int t = 0; t < n_steps; ++t) {
        float rate = __fdiv_rn((float)counts[t], total);   // jnp.mean(do_accept)
        sn += 1;
        float ar_new = __fadd_rn(__fmul_rn(0.9f, ar), __fmul_rn(0.1f, rate));
        if (sn % cfg.stepsize_update_interval == 0) {   // decided with the PRE-update acc_rate
            ss = (ar < cfg.target_acceptance_rate) ? __fdiv_rn(ss, 1.05f) : __fmul_rn(ss, 1.05f);
            ss = fminf(fmaxf(ss, cfg.min_stepsize_scale), cfg.max_stepsize_scale);
        }
        ar = ar_new;
    }
    stepsize[0] = ss; acc_rate[0] = ar; step_nr[0] = sn;
}

int launch_controller(const dpe_mcmc_state *st, const int32_t *counts, int n_steps, int64_t n_total,
                      const dpe_mcmc_config &cfg, cudaStream_t s) {
    k_controller<<<1, 32, 0, s>>>(st->stepsize_dev, st->step_nr_dev, st->acc_rate_dev, counts, n_steps, (float)n_total, cfg);
    return check_cuda(cudaGetLastError(), "k_controller");
}

// ------------------------------------------------------------------------------------------------
// E_loc statistics (loss_function.py): single block reductions, NaN-tolerant (jnp.nanmean)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float *red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(1024) k_moments1(const float *__restrict__ e, int n, const float *__restrict__ cw,
                                                    int clip_mode, float *__restrict__ ec, float *__restrict__ out) {
    __shared__ float red[32];
    const float c = cw[0], w = cw[1];
    float s = 0.f, sc = 0.f, cnt = 0.f, cntc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = e[i];
        float vc = clip_mode == 1 ? fminf(fmaxf(v, c - w), c + w) : c + tanhf((v - c) / w) * w;
        if (v != v) vc = v;
        ec[i] = vc;
        if (v == v) { s += v; cnt += 1.f; }
        if (vc == vc) { sc += vc; cntc += 1.f; }
    }
    s = block_sum(s, red); cnt = block_sum(cnt, red); sc = block_sum(sc, red); cntc = block_sum(cntc, red);
    if (threadIdx.x == 0) { out[0] = s / cnt; out[1] = sc / cntc; }
}

__global__ void __launch_bounds__(1024) k_moments2(const float *__restrict__ e, const float *__restrict__ ec, int n,
                                                    const float *__restrict__ means, float *__restrict__ out) {
    __shared__ float red[32];
    const float m0 = means[0], m1 = means[1];
    float s = 0.f, sc = 0.f, cnt = 0.f, cntc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = e[i] - m0, vc = ec[i] - m1;
        if (v == v) { s = fmaf(v, v, s); cnt += 1.f; }
        if (vc == vc) { sc = fmaf(vc, vc, sc); cntc += 1.f; }
    }
    s = block_sum(s, red); cnt = block_sum(cnt, red); sc = block_sum(sc, red); cntc = block_sum(cntc, red);
    if (threadIdx.x == 0) { out[0] = s / cnt; out[1] = sc / cntc; }
}

// jnp.nanmedian (loss_function.py:20, clipping.center = "median"): radix select on the order-preserving integer image of the
// floats, one block; an even count averages the two middle values.
__device__ __forceinline__ uint32_t f2key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(1024) k_median(const float *__restrict__ e, int n, float *__restrict__ out) {
    __shared__ float red[32];
    __shared__ unsigned hist[256];
    __shared__ uint32_t sel_prefix;
    __shared__ unsigned sel_rank;
    float c = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += (e[i] == e[i]) ? 1.f : 0.f;
    const int cnt = (int)block_sum(c, red);
    if (cnt == 0) { if (threadIdx.x == 0) out[0] = __uint_as_float(0x7fc00000u); return; }
    float vals[2];
    for (int which = 0; which < 2; ++which) {
        if (threadIdx.x == 0) { sel_prefix = 0u; sel_rank = (unsigned)(which ? cnt / 2 : (cnt - 1) / 2); }
        for (int pass = 3; pass >= 0; --pass) {
            for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
            __syncthreads();
            const uint32_t prefix = sel_prefix;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const float v = e[i];
                if (v == v) {
                    const uint32_t k = f2key(v);
                    if (pass == 3 || (k >> (8 * (pass + 1))) == prefix) atomicAdd(&hist[(k >> (8 * pass)) & 255u], 1u);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned r = sel_rank, b = 0;
                while (r >= hist[b]) { r -= hist[b]; ++b; }
                sel_rank = r;
                sel_prefix = (prefix << 8) | b;
            }
            __syncthreads();
        }
        vals[which] = key2f(sel_prefix);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = vals[0] + (vals[1] - vals[0]) * 0.5f;      // linear interpolation, as jnp.nanquantile does
}

// width of the clipping window (loss_function.py:22-27): nanmean((E - center)^2) (metric 0, "std") or nanmean(|E - center|) (1, "mae")
__global__ void __launch_bounds__(1024) k_width(const float *__restrict__ e, int n, const float *__restrict__ center, int metric,
                                                 float *__restrict__ out) {
    __shared__ float red[32];
    const float c = center[0];
    float s = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = e[i] - c;
        if (v == v) { s = metric ? s + fabsf(v) : fmaf(v, v, s); cnt += 1.f; }
    }
    s = block_sum(s, red); cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) out[0] = s / cnt;
}

__global__ void k_bits(uint32_t k0, uint32_t k1, int n, uint32_t *__restrict__ bits) {
    const int h = (n + 1) / 2;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h) return;
    uint32_t x0 = (uint32_t)p, x1 = (h + p < n) ? (uint32_t)(h + p) : 0u;
    threefry2x32(k0, k1, x0, x1);
    bits[p] = x0;
    if (h + p < n) bits[h + p] = x1;
}

__global__ void k_normal(uint32_t k0, uint32_t k1, int n, float *__restrict__ out) {
    const int h = (n + 1) / 2;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h) return;
    uint32_t x0 = (uint32_t)p, x1 = (h + p < n) ? (uint32_t)(h + p) : 0u;
    threefry2x32(k0, k1, x0, x1);
    out[p] = bits_to_normal(x0);
    if (h + p < n) out[h + p] = bits_to_normal(x1);
}

}  // namespace dpe

extern "C" {

int dpe_energy_moments1(const float *e_loc_dev, int32_t n, const float *clip_center_width_dev, int32_t clip_mode,
                        float *e_clipped_dev, float *out2_dev, void *stream) {
    if (!e_loc_dev || !clip_center_width_dev || !e_clipped_dev || !out2_dev || n <= 0) return dpe::set_error(DPE_ERR_ARG, "energy_moments1: bad argument");
    dpe::k_moments1<<<1, 1024, 0, (cudaStream_t)stream>>>(e_loc_dev, n, clip_center_width_dev, clip_mode, e_clipped_dev, out2_dev);
    return dpe::check_cuda(cudaGetLastError(), "k_moments1");
}

int dpe_energy_moments2(const float *e_loc_dev, const float *e_clipped_dev, int32_t n, const float *means_dev,
                        float *out2_dev, void *stream) {
    if (!e_loc_dev || !e_clipped_dev || !means_dev || !out2_dev || n <= 0) return dpe::set_error(DPE_ERR_ARG, "energy_moments2: bad argument");
    dpe::k_moments2<<<1, 1024, 0, (cudaStream_t)stream>>>(e_loc_dev, e_clipped_dev, n, means_dev, out2_dev);
    return dpe::check_cuda(cudaGetLastError(), "k_moments2");
}

int dpe_energy_median(const float *e_dev, int32_t n, float *out_dev, void *stream) {
    if (!e_dev || !out_dev || n <= 0) return dpe::set_error(DPE_ERR_ARG, "energy_median: bad argument");
    dpe::k_median<<<1, 1024, 0, (cudaStream_t)stream>>>(e_dev, n, out_dev);
    return dpe::check_cuda(cudaGetLastError(), "k_median");
}

int dpe_energy_width(const float *e_dev, int32_t n, const float *center_dev, int32_t metric, float *out_dev, void *stream) {
    if (!e_dev || !center_dev || !out_dev || n <= 0 || metric < 0 || metric > 1) return dpe::set_error(DPE_ERR_ARG, "energy_width: bad argument");
    dpe::k_width<<<1, 1024, 0, (cudaStream_t)stream>>>(e_dev, n, center_dev, metric, out_dev);
    return dpe::check_cuda(cudaGetLastError(), "k_width");
}

int dpe_threefry_mcmc_randoms(const uint32_t *keys_dev, int32_t n_walkers, int32_t n_el, uint32_t *new_keys_dev,
                              float *noise_dev, float *thr_dev, void *stream) {
    if (!keys_dev || !new_keys_dev || !noise_dev || !thr_dev || n_walkers <= 0 || n_el <= 0) return dpe::set_error(DPE_ERR_ARG, "threefry_mcmc_randoms: bad argument");
    int n = 3 * n_el, h = (n + 1) / 2;
    long total = (long)n_walkers * h;
    dpe::k_propose<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nullptr, keys_dev, nullptr, n_walkers, n, nullptr,
                                                                                 noise_dev, thr_dev, new_keys_dev, 0);
    return dpe::check_cuda(cudaGetLastError(), "k_propose(randoms)");
}

int dpe_threefry_bits(const uint32_t *key_host, int32_t n, uint32_t *bits_dev, void *stream) {
    if (!key_host || !bits_dev || n <= 0) return dpe::set_error(DPE_ERR_ARG, "threefry_bits: bad argument");
    int h = (n + 1) / 2;
    dpe::k_bits<<<(h + 255) / 256, 256, 0, (cudaStream_t)stream>>>(key_host[0], key_host[1], n, bits_dev);
    return dpe::check_cuda(cudaGetLastError(), "k_bits");
}

int dpe_threefry_normal(const uint32_t *key_host, int32_t n, float *out_dev, void *stream) {
    if (!key_host || !out_dev || n <= 0) return dpe::set_error(DPE_ERR_ARG, "threefry_normal: bad argument");
    int h = (n + 1) / 2;
    dpe::k_normal<<<(h + 255) / 256, 256, 0, (cudaStream_t)stream>>>(key_host[0], key_host[1], n, out_dev);
    return dpe::check_cuda(cudaGetLastError(), "k_normal");
}

int dpe_mcmc_controller(const dpe_mcmc_state *state, const int32_t *accept_counts_dev, int32_t n_steps,
                        int64_t n_walkers_total, const dpe_mcmc_config *cfg, void *stream) {
    if (!state || !accept_counts_dev || !cfg || n_steps < 0 || n_walkers_total <= 0) return dpe::set_error(DPE_ERR_ARG, "mcmc_controller: bad argument");
    if (n_steps == 0) return DPE_OK;
    return dpe::launch_controller(state, accept_counts_dev, n_steps, n_walkers_total, *cfg, (cudaStream_t)stream);
}

}  // extern "C"
