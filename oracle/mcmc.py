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


# ---- exponential radial initialisation (the reference's default `initialization: exponential`, configuration.py:1013) ----
def _exp_radial_tables():
    from ._exp_radial_grid import EXP_RADIAL_XP
    xp = np.asarray(EXP_RADIAL_XP, np.float64)
    return xp, 1.0 - np.exp(-xp) * (1.0 + xp + 0.5 * xp * xp)


def generate_exp_distributed(rng, batch_shape, k=1.0):
    """utils/utils.py:387-506: points with radial pdf r^2 exp(-k r): direction from a normalised Gaussian, radius from the tabulated
    inverse CDF (jnp.interp) of a uniform sample."""
    xp, Fp = _exp_radial_tables()
    key_uniform, key_gaussian = threefry.split(rng)
    r = threefry.normal(key_gaussian, tuple(batch_shape) + (3,))
    u = threefry.uniform(key_uniform, tuple(batch_shape))
    x = np.interp(u, Fp, xp).astype(f32)
    return (r / np.linalg.norm(r, axis=-1, keepdims=True)).astype(f32) * (x / f32(k))[..., None]


def get_effective_charge(Z, n, s_lower_shell=0.85, s_same_shell=0.35):
    """orbitals.py:854-877 (Slater's rules)."""
    shielding = 0
    for n_shell in range(1, n + 1):
        n_electrons_in_shell = 2 * n_shell ** 2
        if n_shell == n:
            n_el_in_lower_shells = sum(2 * k ** 2 for k in range(1, n_shell))
            n_electrons_in_shell = min(Z - n_el_in_lower_shells, n_electrons_in_shell) - 1
        if n_shell == n:
            shielding += n_electrons_in_shell * s_same_shell
        elif n_shell == n - 1:
            shielding += n_electrons_in_shell * s_lower_shell
        else:
            shielding += n_electrons_in_shell
    return max(Z - shielding, 1)


def get_electron_configuration(n_el):
    """orbitals.py:880-891: electrons per principal quantum number in Madelung order."""
    shells = "1s,2s,2p,3s,3p,4s,3d,4p,5s,4d,5p,6s,4f,5d,6p,5f,6d".split(",")
    cap = dict(s=2, p=6, d=10, f=14)
    cfg = {}
    while n_el > 0:
        shell = shells.pop(0)
        n_in = min(n_el, cap[shell[1]])
        cfg[int(shell[0])] = cfg.get(int(shell[0]), 0) + n_in
        n_el -= n_in
    return cfg


def _initialize_walkers_around_atom(rng, R, Z, n_walkers, n_el, n_up):
    """orbitals.py:894-910."""
    r, spin = [], []
    for n, n_in_shell in get_electron_configuration(n_el).items():
        subkey, rng = threefry.split(rng)
        exponent = 2 * get_effective_charge(Z, n, 1.0, 0.7) / n
        r.append(generate_exp_distributed(subkey, [n_walkers, n_in_shell], exponent))
        n_up_in_shell = max(min(n_in_shell // 2, n_up), n_in_shell - n_el + n_up)
        n_dn_in_shell = n_in_shell - n_up_in_shell
        n_up -= n_up_in_shell
        n_el -= n_in_shell
        spin += [0] * n_up_in_shell + [1] * n_dn_in_shell
    r = np.concatenate(r, axis=1) + np.asarray(R, f32)
    is_up = np.array(spin) == 0
    return r[:, is_up, :], r[:, ~is_up, :]


def initialize_walkers_with_exponential_radial_pdf(rng, R, Z, n_walkers, n_el, n_up, el_ion_mapping):
    """orbitals.py:913-928."""
    assert n_el == len(el_ion_mapping)
    up_map, dn_map = np.array(el_ion_mapping)[:n_up], np.array(el_ion_mapping)[n_up:]
    r_up, r_dn = [], []
    for ind_ion, (R_, Z_) in enumerate(zip(R, Z)):
        subkey, rng = threefry.split(rng)
        n_up_ion = int(sum(up_map == ind_ion))
        n_el_ion = int(sum(dn_map == ind_ion)) + n_up_ion
        if n_el_ion == 0:
            continue
        a_up, a_dn = _initialize_walkers_around_atom(subkey, R_, int(Z_), n_walkers, n_el_ion, n_up_ion)
        r_up.append(a_up)
        r_dn.append(a_dn)
    return np.concatenate(r_up + r_dn, axis=1).astype(f32)


def initialize_around_nuclei(n_walkers, R, Z, el_ion_mapping, seed, init_method="gaussian", n_up=None) -> OracleMCMCState:
    """mcmc.py:39-91 with spin_initialization='el_ion_mapping'."""
    rng = threefry.prng_key(seed)
    rng_r, _rng_spin, rng = threefry.split(rng, 3)
    n_el = len(el_ion_mapping)
    if init_method == "gaussian":
        r0 = threefry.normal(rng_r, (n_walkers, n_el, 3)) + np.asarray(R, f32)[np.asarray(el_ion_mapping)]
    else:
        r0 = initialize_walkers_with_exponential_radial_pdf(rng_r, R, Z, n_walkers, n_el, n_up, el_ion_mapping)
    return OracleMCMCState(
        r=r0.astype(f32), R=np.asarray(R, f32), Z=np.asarray(Z, np.int32),
        log_psi_sqr=-np.ones(n_walkers, f32) * f32(1000), walker_age=np.zeros(n_walkers, np.int32),
        rng_state=threefry.split(rng, n_walkers))


def _el_ion(r, R):
    """utils/utils.py get_el_ion_distance_matrix: diff[b, i, J] = r_i - R_J, dist = |diff| (float32)."""
    diff = (r[:, :, None, :] - R[None, None, :, :]).astype(f32)
    return diff, np.sqrt(np.sum(diff * diff, axis=-1, dtype=f32)).astype(f32)


def make_mcmc_step(func: Callable[[np.ndarray], np.ndarray], state: OracleMCMCState, *, max_age=20,
                   stepsize_update_interval=100, target_acceptance_rate=0.5, min_stepsize_scale=1e-2,
                   max_stepsize_scale=1.0, allreduce_mean=lambda x: x, return_mask=False, proposal="normal", proposal_kw=None):
    """mcmc.py:345-387 with every proposal of mcmc.py:175-284 (`proposal_kw`: r_min, r_max, langevin_scale of the local / langevin ones).
    func(r[B,N,3] f32) -> log_psi_sqr[B] f32."""
    B, N, _ = state.r.shape
    proposal_kw = proposal_kw or {}
    new_keys, noise, thr = threefry.mcmc_step_randoms(state.rng_state, N, proposal)   # same subkey for noise and threshold
    log_q_ratio = np.zeros(B, f32)
    ss = f32(state.stepsize)
    if proposal == "normal_one_el":          # mcmc.py:183-193: only electron step_nr % n_el moves
        r_new = state.r.copy()
        idx = int(state.step_nr) % N
        r_new[:, idx, :] = (state.r[:, idx, :] + noise * ss).astype(f32)
    elif proposal in ("local", "local_one_el"):     # mcmc.py:212-253: step size proportional to the distance to the closest nucleus
        r_min, r_max = f32(proposal_kw.get("r_min", 0.1)), f32(proposal_kw.get("r_max", 1.0))
        closest = lambda rr: np.min(_el_ion(rr, state.R)[1], axis=-1)
        if proposal == "local":
            s = (ss * np.clip(closest(state.r), r_min, r_max)).astype(f32)                       # [B, N]
            r_new = (state.r + noise * s[..., None]).astype(f32)
            s_new = (ss * np.clip(closest(r_new), r_min, r_max)).astype(f32)
            dist_sqr = np.sum((r_new - state.r) ** 2, axis=-1, dtype=f32)
            lq = f32(3) * (np.log(s) - np.log(s_new)) + f32(0.5) * dist_sqr * (f32(1) / s ** 2 - f32(1) / s_new ** 2)
            log_q_ratio = np.sum(lq, axis=-1, dtype=f32)
        else:
            idx = int(state.step_nr) % N
            s = (ss * np.clip(closest(state.r)[:, idx], r_min, r_max)).astype(f32)               # [B]
            r_new = state.r.copy()
            r_new[:, idx, :] = (state.r[:, idx, :] + noise * s[:, None]).astype(f32)
            s_new = (ss * np.clip(closest(r_new)[:, idx], r_min, r_max)).astype(f32)
            dist_sqr = np.sum((r_new[:, idx, :] - state.r[:, idx, :]) ** 2, axis=-1, dtype=f32)
            log_q_ratio = (f32(3) * (np.log(s) - np.log(s_new)) + f32(0.5) * dist_sqr * (f32(1) / s ** 2 - f32(1) / s_new ** 2)).astype(f32)
    elif proposal == "langevin":             # mcmc.py:205-209, 256-284: local step size + drift towards the nuclei
        scale = f32(proposal_kw.get("langevin_scale", 1.0))
        r_min, r_max = f32(proposal_kw.get("r_min", 0.2)), f32(proposal_kw.get("r_max", 2.0))

        def step_and_bias(rr):
            diff, dist = _el_ion(rr, state.R)
            g = (-scale * np.sum(diff * state.Z.astype(f32)[:, None] / dist[..., None], axis=-2, dtype=f32)).astype(f32)   # [B, N, 3]
            return (ss * np.clip(np.min(dist, axis=-1, keepdims=True), r_min, r_max)).astype(f32), g                     # [B, N, 1]

        s, g = step_and_bias(state.r)
        r_new = (state.r + noise * s + g * s ** 2).astype(f32)
        s_new, g_new = step_and_bias(r_new)
        d_fwd = np.sum((r_new - state.r - g * s ** 2) ** 2, axis=-1, dtype=f32)
        d_rev = np.sum((state.r - r_new - g_new * s_new ** 2) ** 2, axis=-1, dtype=f32)
        s1, s1n = s[..., 0], s_new[..., 0]
        lq = f32(3) * (np.log(s1) - np.log(s1n)) + f32(0.5) * (d_fwd / s1 ** 2 - d_rev / s1n ** 2)
        log_q_ratio = np.sum(lq, axis=-1, dtype=f32)
    else:                                    # normal (mcmc.py:175-180) / cauchy (:196-201); log_q_ratio = 0
        r_new = (state.r + noise * ss).astype(f32)
    lp_new = np.asarray(func(r_new), f32)
    with np.errstate(over="ignore"):
        p_accept = np.exp(((lp_new - state.log_psi_sqr).astype(f32) + log_q_ratio).astype(f32)).astype(f32)
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
