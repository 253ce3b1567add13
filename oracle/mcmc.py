"""ORACLE (test infrastructure, never the product path): numpy restatement of DeepErwin's
Metropolis-Hastings walker update (src/deeperwin/mcmc.py) and of the energy statistics / clipping of
src/deeperwin/optimization/loss_function.py:12-109.

PARITY UNPINNED for the float parts (no JAX in this image, no golden vectors in the reference's tests);
the RNG stream is pinned (oracle/threefry.py).  Integer bookkeeping (keys, accept masks given the same
log psi^2, ages, step_nr) is exact integer arithmetic and is what the CUDA path must match bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Callable

import numpy as np

from . import threefry

f32 = np.float32


@dataclass
class OracleMCMCState:
    """mcmc.py:20-33 MCMCState fields/dtypes."""
    r: np.ndarray              # f32 [B,N,3]
    R: np.ndarray              # f32 [I,3]
    Z: np.ndarray              # int [I]
    log_psi_sqr: np.ndarray    # f32 [B]
    walker_age: np.ndarray     # i32 [B]
    rng_state: np.ndarray      # u32 [B,2]
    stepsize: np.float32 = f32(1e-2)
    step_nr: int = 0
    acc_rate: np.float32 = f32(0.0)


def initialize_around_nuclei(n_walkers, R, Z, el_ion_mapping, seed) -> OracleMCMCState:
    """mcmc.py:39-91 with init_method='gaussian', spin_initialization='el_ion_mapping'."""
    rng = threefry.prng_key(seed)
    rng_r, _rng_spin, rng = threefry.split(rng, 3)
    n_el = len(el_ion_mapping)
    r0 = threefry.normal(rng_r, (n_walkers, n_el, 3)) + np.asarray(R, f32)[np.asarray(el_ion_mapping)]
    return OracleMCMCState(
        r=r0.astype(f32), R=np.asarray(R, f32), Z=np.asarray(Z, np.int32),
        log_psi_sqr=-np.ones(n_walkers, f32) * f32(1000), walker_age=np.zeros(n_walkers, np.int32),
        rng_state=threefry.split(rng, n_walkers))


def make_mcmc_step(func: Callable[[np.ndarray], np.ndarray], state: OracleMCMCState, *, max_age=20,
                   stepsize_update_interval=100, target_acceptance_rate=0.5, min_stepsize_scale=1e-2,
                   max_stepsize_scale=1.0, allreduce_mean=lambda x: x, return_mask=False):
    """mcmc.py:345-387 with the `normal` proposal (mcmc.py:175-180).
    func(r[B,N,3] f32) -> log_psi_sqr[B] f32."""
    B, N, _ = state.r.shape
    new_keys, noise, thr = threefry.mcmc_step_randoms(state.rng_state, N)   # same subkey for noise and threshold
    r_new = (state.r + noise * f32(state.stepsize)).astype(f32)
    lp_new = np.asarray(func(r_new), f32)
    with np.errstate(over="ignore"):
        p_accept = np.exp((lp_new - state.log_psi_sqr).astype(f32)).astype(f32)
    do_accept = (p_accept > thr) | (state.walker_age >= max_age)
    age = np.where(do_accept, 0, state.walker_age + 1).astype(np.int32)
    lp = np.where(do_accept, lp_new, state.log_psi_sqr).astype(f32)
    r = np.where(do_accept[:, None, None], r_new, state.r).astype(f32)
    acceptance_rate = f32(allreduce_mean(f32(np.mean(do_accept, dtype=np.float32))))
    step_nr = state.step_nr + 1
    acc_rate = f32(f32(0.9) * state.acc_rate + f32(0.1) * acceptance_rate)
    stepsize = f32(state.stepsize)
    if step_nr % stepsize_update_interval == 0:                              # uses the PRE-update acc_rate
        stepsize = f32(stepsize / f32(1.05)) if state.acc_rate < f32(target_acceptance_rate) else f32(stepsize * f32(1.05))
        stepsize = f32(np.clip(stepsize, f32(min_stepsize_scale), f32(max_stepsize_scale)))
    new = replace(state, r=r, log_psi_sqr=lp, walker_age=age, rng_state=new_keys, stepsize=stepsize,
                  step_nr=step_nr, acc_rate=acc_rate)
    return (new, do_accept) if return_mask else new


def run_mcmc_steps(func, state: OracleMCMCState, n_steps, **cfg) -> OracleMCMCState:
    """mcmc.py:389-406: recompute log_psi_sqr with the current params, then n_steps Metropolis steps."""
    state = replace(state, log_psi_sqr=np.asarray(func(state.r), f32))
    for _ in range(n_steps):
        state = make_mcmc_step(func, state, **cfg)
    return state


# ----------------------------------------------------------------------------- loss_function.py
def init_clipping_state():
    """loss_function.py:12-16."""
    return f32(0.0), f32(1e12)


def _center_and_width(E, clip_by=5.0, center="mean", width_metric="std", allreduce_mean=lambda x: x):
    """loss_function.py:19-30."""
    c = f32(allreduce_mean(f32(np.nanmean(E) if center == "mean" else np.nanmedian(E))))
    if width_metric == "mae":
        w = f32(allreduce_mean(f32(np.nanmean(np.abs(E - c)))))
    else:
        w = f32(np.sqrt(f32(allreduce_mean(f32(np.nanmean((E - c) ** 2))))))
    return c, f32(w * f32(clip_by))


def energy_statistics(E_loc, clipping_state, *, name="tanh", clip_by=5.0, center="mean", width_metric="std",
                      from_previous_step=True, allreduce_mean=lambda x: x):
    """total_energy (loss_function.py:89-109) without the gradient: returns (loss, new_clipping_state, aux)."""
    E = np.asarray(E_loc, f32)
    kw = dict(clip_by=clip_by, center=center, width_metric=width_metric, allreduce_mean=allreduce_mean)
    E_mean = f32(allreduce_mean(f32(np.nanmean(E))))
    E_var = f32(allreduce_mean(f32(np.nanmean((E - E_mean) ** 2))))
    c, w = clipping_state
    if not from_previous_step or c is None:
        c, w = _center_and_width(E, **kw)
    if name == "hard":
        Ec = np.clip(E, c - w, c + w).astype(f32)
    else:
        Ec = (c + np.tanh((E - c) / w) * w).astype(f32)
    new_state = _center_and_width(Ec, **kw)
    Ec_mean = f32(allreduce_mean(f32(np.nanmean(Ec))))
    Ec_var = f32(allreduce_mean(f32(np.nanmean((Ec - Ec_mean) ** 2))))
    aux = dict(E_mean=E_mean, E_var=E_var, E_mean_clipped=Ec_mean, E_var_clipped=Ec_var, E_loc_clipped=Ec, E_loc=E)
    return Ec_mean, new_state, aux
