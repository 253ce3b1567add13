"""Drop-in for src/deeperwin/mcmc.py: MCMCState (:20-146) and MetropolisHastingsMonteCarlo (:314-412)
with the default `normal` proposal (:175-180).  Walker state lives in torch CUDA tensors with the
reference's field names and dtypes; every step runs in libdpe_b200.so (fused threefry proposal,
forward pass, accept/reject, step-size controller).

One process per GPU: `split_across_devices` keeps this rank's contiguous block of walkers
(mcmc.py:105-129 with n_local_devices = 1), `merge_devices` all-gathers them (mcmc.py:131-146)."""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass, field, replace
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, utils
from ._lib import DpeMcmcConfig, DpeMcmcState
from .configuration import MCMCConfig, PhysicalConfig


def PRNGKey(seed: int) -> torch.Tensor:
    """jax.random.PRNGKey: uint32[2] = [hi32, lo32] (host tensor)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return torch.tensor([seed >> 32, seed & 0xFFFFFFFF], dtype=torch.int64).to(torch.uint32)


def _key_array(key) -> "C.Array":
    k = np.asarray(key.cpu() if isinstance(key, torch.Tensor) else key).astype(np.uint32).reshape(2)
    return (C.c_uint32 * 2)(int(k[0]), int(k[1]))


def random_bits(key, n: int, device) -> torch.Tensor:
    """jax `random_bits(key, 32, (n,))` on the GPU."""
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.dpe_threefry_bits(_key_array(key), n, C.c_void_p(out.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "dpe_threefry_bits")
    return out.view(torch.uint32)


def split(key, num: int = 2, device="cuda") -> torch.Tensor:
    """jax.random.split(key, num) -> uint32[num, 2]."""
    return random_bits(key, 2 * num, device).reshape(num, 2)


def normal(key, shape, device="cuda") -> torch.Tensor:
    """jax.random.normal(key, shape, float32)."""
    lib = _lib.load()
    n = int(np.prod(shape))
    out = torch.empty(n, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.dpe_threefry_normal(_key_array(key), n, C.c_void_p(out.data_ptr()),
                                           C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "dpe_threefry_normal")
    return out.reshape(*shape)


@dataclass
class MCMCState:
    """mcmc.py:20-33."""
    r: torch.Tensor                      # f32 [B, n_el, 3]
    R: torch.Tensor                      # f32 [n_ions, 3]
    Z: torch.Tensor                      # int32 [n_ions]
    log_psi_sqr: Optional[torch.Tensor] = None   # f32 [B]
    walker_age: Optional[torch.Tensor] = None    # int32 [B]
    rng_state: Optional[torch.Tensor] = None     # uint32 [B, 2]
    stepsize: Optional[torch.Tensor] = None      # f32 scalar (init 1e-2)
    step_nr: Optional[torch.Tensor] = None       # int32 scalar
    acc_rate: Optional[torch.Tensor] = None      # f32 scalar
    _step_nr_host: Optional[int] = field(default=None, repr=False, compare=False)

    def __post_init__(self):
        dev = self.r.device
        if self.stepsize is None:
            self.stepsize = torch.full((), 1e-2, dtype=torch.float32, device=dev)
        if self.step_nr is None:
            self.step_nr = torch.zeros((), dtype=torch.int32, device=dev)
            self._step_nr_host = 0
        if self.acc_rate is None:
            self.acc_rate = torch.zeros((), dtype=torch.float32, device=dev)

    def build_batch(self, fixed_params: Dict):
        """mcmc.py:36-37."""
        return self.r, self.R, self.Z, fixed_params

    @classmethod
    def initialize_around_nuclei(cls, n_walkers, physical_config: PhysicalConfig, init_method, spin_initialization, rng,
                                 device="cuda"):
        """mcmc.py:39-91 for spin_initialization='el_ion_mapping': init_method 'gaussian' (unit normal around the mapped nucleus)
        or 'exponential' (the reference's default: per-shell exponential radial pdf with Slater-rule exponents,
        orbitals.py:854-928, utils/utils.py:387-506).  Setup-time host logic on top of the library's threefry streams."""
        if spin_initialization != "el_ion_mapping":
            raise NotImplementedError(f"Unknown spin initialization: {spin_initialization}")
        if init_method not in ("gaussian", "exponential"):
            raise NotImplementedError(f"Unknown initialization: {init_method}")
        device = torch.device(device)
        keys = split(rng, 3, device).cpu()
        rng_r, rng = keys[0], keys[2]
        n_el = physical_config.n_electrons
        R = torch.tensor(physical_config.R, dtype=torch.float32, device=device)
        if init_method == "gaussian":
            r0 = normal(rng_r, (n_walkers, n_el, 3), device)
            r0 = r0 + R[torch.tensor(physical_config.el_ion_mapping, dtype=torch.long, device=device)]
        else:
            r0 = initialize_walkers_with_exponential_radial_pdf(rng_r, physical_config.R, physical_config.Z, n_walkers, n_el,
                                                                 physical_config.n_up, physical_config.el_ion_mapping, device)
        return cls(r=r0, R=R, Z=torch.tensor(physical_config.Z, dtype=torch.int32, device=device),
                   log_psi_sqr=-torch.ones(n_walkers, dtype=torch.float32, device=device) * 1000,
                   walker_age=torch.zeros(n_walkers, dtype=torch.int32, device=device),
                   rng_state=split(rng, n_walkers, device))

    @classmethod
    def resize_or_init(cls, mcmc_state, mcmc_config: MCMCConfig, physical_config: PhysicalConfig, rng, device="cuda"):
        """mcmc.py:93-103."""
        if mcmc_state:
            if mcmc_state.r.ndim == 4:
                mcmc_state = mcmc_state.merge_devices()
            return resize_nr_of_walkers(mcmc_state, mcmc_config.n_walkers)
        return cls.initialize_around_nuclei(mcmc_config.n_walkers, physical_config, mcmc_config.initialization,
                                            mcmc_config.spin_initialization, rng, device)

    def split_across_devices(self):
        """mcmc.py:105-129: contiguous blocks of B/n_dev walkers; this process keeps block `rank` with a
        leading local-device axis of size 1, the unbatched fields are tiled to [1, ...]."""
        assert self.r.ndim == 3, "State is already split across devices"
        n_dev, rk = utils.world_size(), utils.rank()
        assert len(self.r) % n_dev == 0, f"Number of samples ({len(self.r)}) is not evenly divisible across devices ({n_dev})"
        n = len(self.r) // n_dev

        def _split(x):
            return x[rk * n:(rk + 1) * n][None].contiguous()

        def _tile(x):
            return x[None].clone()

        return MCMCState(r=_split(self.r), log_psi_sqr=_split(self.log_psi_sqr), walker_age=_split(self.walker_age),
                         rng_state=_split(self.rng_state), R=_tile(self.R), Z=_tile(self.Z), stepsize=_tile(self.stepsize),
                         step_nr=_tile(self.step_nr), acc_rate=_tile(self.acc_rate), _step_nr_host=self._step_nr_host)

    def merge_devices(self):
        """mcmc.py:131-146 (real all-gather instead of the psum-of-padded emulation, utils.py:100-112)."""
        if self.r.ndim == 3:
            return self
        assert self.r.ndim == 4, "State is not split across devices"
        g = lambda x: utils.all_gather_batch(x[0])
        return MCMCState(r=g(self.r), log_psi_sqr=g(self.log_psi_sqr), walker_age=g(self.walker_age), rng_state=g(self.rng_state),
                         R=self.R[0], Z=self.Z[0], stepsize=self.stepsize[0], step_nr=self.step_nr[0], acc_rate=self.acc_rate[0],
                         _step_nr_host=self._step_nr_host)


# ---- exponential radial initialisation (setup time; orbitals.py:854-928, utils/utils.py:387-506) ---------------------------
def uniform(key, shape, device="cuda") -> torch.Tensor:
    """jax.random.uniform(key, shape) in [0, 1): ((bits >> 9) | 0x3F800000).view(f32) - 1 on the library's threefry bits."""
    n = int(np.prod(shape)) if len(shape) else 1
    bits = random_bits(key, n, device).cpu().numpy().astype(np.uint32)
    u = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return torch.from_numpy(np.maximum(np.float32(0.0), u).reshape(shape)).to(device)


def generate_exp_distributed(rng, batch_shape, k=1.0, device="cuda") -> torch.Tensor:
    """utils/utils.py:387-506: radial pdf r^2 exp(-k r); radius by the tabulated inverse CDF (jnp.interp), direction from a
    normalised Gaussian.  The table ordinates are the closed-form CDF (see _exp_radial_grid.py)."""
    from ._exp_radial_grid import EXP_RADIAL_XP
    xp = np.asarray(EXP_RADIAL_XP, np.float64)
    Fp = 1.0 - np.exp(-xp) * (1.0 + xp + 0.5 * xp * xp)
    keys = split(rng, 2, device).cpu()
    key_uniform, key_gaussian = keys[0], keys[1]
    r = normal(key_gaussian, tuple(batch_shape) + (3,), device)
    u = uniform(key_uniform, tuple(batch_shape), device).cpu().numpy()
    x = torch.from_numpy(np.interp(u, Fp, xp).astype(np.float32)).to(device)
    return r / r.norm(dim=-1, keepdim=True) * (x / np.float32(k))[..., None]


def _get_effective_charge(Z: int, n: int, s_lower_shell=0.85, s_same_shell=0.35):
    """orbitals.py:854-877: Slater's rules."""
    shielding = 0
    for n_shell in range(1, n + 1):
        n_in_shell = 2 * n_shell ** 2
        if n_shell == n:
            n_in_shell = min(Z - sum(2 * k ** 2 for k in range(1, n_shell)), n_in_shell) - 1
            shielding += n_in_shell * s_same_shell
        elif n_shell == n - 1:
            shielding += n_in_shell * s_lower_shell
        else:
            shielding += n_in_shell
    return max(Z - shielding, 1)


def _get_electron_configuration(n_el: int):
    """orbitals.py:880-891: electrons per principal quantum number, shells filled in Madelung order."""
    order = "1s,2s,2p,3s,3p,4s,3d,4p,5s,4d,5p,6s,4f,5d,6p,5f,6d".split(",")
    capacity = dict(s=2, p=6, d=10, f=14)
    config: Dict[int, int] = {}
    for shell in order:
        if n_el <= 0:
            break
        n_here = min(n_el, capacity[shell[1]])
        config[int(shell[0])] = config.get(int(shell[0]), 0) + n_here
        n_el -= n_here
    return config


def initialize_walkers_with_exponential_radial_pdf(rng, R, Z, n_walkers, n_el, n_up, el_ion_mapping, device="cuda"):
    """orbitals.py:894-928: per ion (one subkey each), per shell (one subkey each) exponential clouds with exponent 2 Z_eff / n;
    spin-up electrons of all ions first, then spin-down."""
    if n_el != len(el_ion_mapping):
        raise ValueError("Number of electrons does not match the number of indices in el_ion_mapping")
    mapping = np.asarray(el_ion_mapping)
    r_up, r_dn = [], []
    for ind_ion, (R_, Z_) in enumerate(zip(R, Z)):
        ks = split(rng, 2, device).cpu()
        subkey, rng = ks[0], ks[1]
        n_up_ion = int((mapping[:n_up] == ind_ion).sum())
        n_el_ion = int((mapping[n_up:] == ind_ion).sum()) + n_up_ion
        if n_el_ion == 0:
            continue
        clouds, spin = [], []
        n_el_left, n_up_left = n_el_ion, n_up_ion
        for n, n_in_shell in _get_electron_configuration(n_el_ion).items():
            ks = split(subkey, 2, device).cpu()
            shell_key, subkey = ks[0], ks[1]
            exponent = 2 * _get_effective_charge(int(Z_), n, 1.0, 0.7) / n
            clouds.append(generate_exp_distributed(shell_key, [n_walkers, n_in_shell], exponent, device))
            n_up_in_shell = max(min(n_in_shell // 2, n_up_left), n_in_shell - n_el_left + n_up_left)
            spin += [0] * n_up_in_shell + [1] * (n_in_shell - n_up_in_shell)
            n_up_left -= n_up_in_shell
            n_el_left -= n_in_shell
        r_atom = torch.cat(clouds, dim=1) + torch.tensor(R_, dtype=torch.float32, device=device)
        is_up = torch.tensor(spin, device=device) == 0
        r_up.append(r_atom[:, is_up, :])
        r_dn.append(r_atom[:, ~is_up, :])
    return torch.cat(r_up + r_dn, dim=1).contiguous()


def _resize_array(x, new_length):
    """mcmc.py:154-161."""
    old_length = x.shape[0]
    if new_length < old_length:
        return x[:new_length]
    n_replicas, n_remainder = new_length // old_length, new_length % old_length
    return torch.cat([x for _ in range(n_replicas)] + [x[:n_remainder]], dim=0)


def resize_nr_of_walkers(state: MCMCState, n_walkers_new):
    """mcmc.py:164-172."""
    if n_walkers_new == len(state.r):
        return state
    return replace(state, r=_resize_array(state.r, n_walkers_new), log_psi_sqr=_resize_array(state.log_psi_sqr, n_walkers_new),
                   walker_age=_resize_array(state.walker_age, n_walkers_new),
                   rng_state=split(state.rng_state[0], n_walkers_new, state.r.device))


def plan_segments(step_nr: int, n_steps: int, interval: int):
    """Multi-GPU schedule of one run_mcmc_steps call: the step size only changes when
    step_nr % stepsize_update_interval == 0 (mcmc.py:372-377), so the per-step scalar all-reduce of the
    reference (mcmc.py:367) is deferred to the end of each segment that ends on such a boundary (the
    acceptance-rate EMA is replayed exactly from the all-reduced integer counts)."""
    segs, done = [], 0
    while done < n_steps:
        seg = min(n_steps - done, interval - (step_nr + done) % interval)
        segs.append(seg)
        done += seg
    return segs


class MetropolisHastingsMonteCarlo:
    """mcmc.py:314-412. Holds the MCMC logic and configuration, not the state."""

    def __init__(self, mcmc_config: MCMCConfig):
        self.config = mcmc_config
        # mcmc.py:330-343; 0-2 have log_q_ratio = 0, 3-5 carry it into the acceptance probability
        proposals = {"normal": 0, "cauchy": 1, "normal_one_el": 2, "local": 3, "local_one_el": 4, "langevin": 5}
        prop = self.config.proposal
        if prop.name not in proposals:
            raise NotImplementedError("Unknown MCMC proposal type")   # mcmc.py:343
        self._cfg = DpeMcmcConfig(int(mcmc_config.max_age), int(mcmc_config.stepsize_update_interval),
                                  float(mcmc_config.target_acceptance_rate), float(mcmc_config.min_stepsize_scale),
                                  float(mcmc_config.max_stepsize_scale), proposals[prop.name],
                                  float(getattr(prop, "r_min", 0.0)), float(getattr(prop, "r_max", 0.0)), float(getattr(prop, "langevin_scale", 0.0)))
        self.last_accept_counts: Optional[torch.Tensor] = None
        self._work = {}            # persistent step buffers per (walker shape, device, n_steps)

    def _run_mcmc_steps(self, func, state: MCMCState, params, n_up, n_dn, fixed_params, n_steps) -> MCMCState:
        """mcmc.py:389-406. `func` must be the log_psi_sqr callable of build_log_psi_squared."""
        engine = getattr(func, "engine", None)
        if engine is None:
            raise TypeError("run_mcmc_steps needs the log_psi_sqr callable returned by deeperwin_b200.build_log_psi_squared")
        split_axis = state.r.ndim == 4
        sq = (lambda x: x[0]) if split_axis else (lambda x: x)
        engine.set_params(params)
        engine.set_geometry(sq(state.R), sq(state.Z))
        engine.set_tao_cache(((fixed_params or {}).get("cache") or {}).get("taos"))
        # functional update: inputs are never mutated (SURVEY.md 8b conventions).  The steps run on PERSISTENT work buffers of this object
        # (copy in, advance in place, copy out): the library sees the same pointers on every call of a run, so its CUDA-graph replay
        # of repeated calls (dpe_mcmc_steps) applies to the public API as well, whatever the allocator does with the state tensors.
        src = [sq(state.r).to(torch.float32), sq(state.log_psi_sqr).to(torch.float32), sq(state.walker_age).to(torch.int32),
               sq(state.rng_state).contiguous().view(torch.int32), sq(state.stepsize).to(torch.float32).reshape(1), sq(state.step_nr).to(torch.int32).reshape(1),
               sq(state.acc_rate).to(torch.float32).reshape(1)]
        wkey = (tuple(src[0].shape), src[0].device, max(n_steps, 1))
        work = self._work.get(wkey)
        if work is None:
            if len(self._work) >= 4:
                self._work.pop(next(iter(self._work)))
            work = self._work[wkey] = [torch.empty_like(t, memory_format=torch.contiguous_format) for t in src] + \
                                      [torch.zeros(max(n_steps, 1), dtype=torch.int32, device=src[0].device)]
        torch._foreach_copy_(work[:7], src)
        r, lp, age, keys, stepsize, step_nr, acc_rate, counts = work
        B = r.shape[0]
        st = DpeMcmcState(r.data_ptr(), lp.data_ptr(), age.data_ptr(), keys.data_ptr(), stepsize.data_ptr(), step_nr.data_ptr(),
                          acc_rate.data_ptr())
        ws_n = utils.world_size()
        step_host = state._step_nr_host
        if ws_n == 1:
            engine.mcmc_steps(st, B, n_steps, self._cfg, True, True, counts)
        else:
            if step_host is None:
                step_host = int(step_nr.item())
            if n_steps == 0:
                engine.mcmc_steps(st, B, 0, self._cfg, True, False, counts)
            done = 0
            for seg in plan_segments(step_host, n_steps, self._cfg.stepsize_update_interval):
                seg_counts = counts[done:done + seg]
                engine.mcmc_steps(st, B, seg, self._cfg, done == 0, False, seg_counts)
                torch.distributed.all_reduce(seg_counts)     # acceptance rate: pmean(mean(do_accept)), mcmc.py:367
                engine.mcmc_controller(st, seg_counts, seg, B * ws_n, self._cfg)
                done += seg
        out = [torch.empty_like(t) for t in work]
        torch._foreach_copy_(out, work)
        r, lp, age, keys, stepsize, step_nr, acc_rate, counts = out
        keys = keys.view(torch.uint32)
        self.last_accept_counts = counts[:n_steps]
        known = state._step_nr_host if state._step_nr_host is not None else step_host
        un = (lambda x: x[None]) if split_axis else (lambda x: x)
        return MCMCState(r=un(r), R=state.R, Z=state.Z, log_psi_sqr=un(lp), walker_age=un(age), rng_state=un(keys),
                         stepsize=un(stepsize.reshape(())), step_nr=un(step_nr.reshape(())), acc_rate=un(acc_rate.reshape(())),
                         _step_nr_host=None if known is None else known + n_steps)

    # the reference wraps _run_mcmc_steps in jax.pmap (mcmc.py:325-327); one process per GPU needs no wrapper
    def run_mcmc_steps(self, func, state, params, n_up, n_dn, fixed_params, n_steps):
        return self._run_mcmc_steps(func, state, params, n_up, n_dn, fixed_params, n_steps)

    def run_inter_steps(self, func, state: MCMCState, params, n_up, n_dn, fixed_params):
        """mcmc.py:408-409."""
        return self.run_mcmc_steps(func, state, params, n_up, n_dn, fixed_params, self.config.n_inter_steps)

    def run_burn_in(self, func, state: MCMCState, params, n_up, n_dn, fixed_params):
        """mcmc.py:411-412."""
        return self.run_mcmc_steps(func, state, params, n_up, n_dn, fixed_params, self.config.n_burn_in)
