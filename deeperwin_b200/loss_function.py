"""Energy statistics and clipping of src/deeperwin/optimization/loss_function.py:12-109 (the forward part of
`total_energy`): local reductions in libdpe_b200.so, cross-GPU `pmean` as two 2-float NCCL all-reduces.
The parameter gradient (loss_function.py:111-154) is a "next" row of the scope table."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, utils
from .configuration import ClippingConfig


def init_clipping_state(is_complex=False, device="cuda"):
    """loss_function.py:12-16."""
    if is_complex:
        raise NotImplementedError("complex wavefunctions are outside the hot-path scope")
    return torch.tensor(0.0, dtype=torch.float32, device=device), torch.tensor(1e12, dtype=torch.float32, device=device)


def build_total_energy(get_local_energy, clipping_config: ClippingConfig):
    """Returns total_energy(params, clipping_state, spin_state, batch) -> (loss, (clipping_state, aux)) with the
    aux keys of loss_function.py:100-107."""
    if not clipping_config.from_previous_step:
        raise NotImplementedError("clipping.from_previous_step=False")
    clip_mode = {"tanh": 0, "hard": 1}[clipping_config.name]
    lib = _lib.load()

    def total_energy(params, state, spin_state, batch):
        r, R, Z, fixed_params = batch
        E_loc = get_local_energy(params, spin_state, r, R, Z, fixed_params).reshape(-1).contiguous()
        dev = E_loc.device
        n = E_loc.numel()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        cw = torch.stack([state[0].reshape(()), state[1].reshape(())]).to(torch.float32).contiguous()
        E_clipped = torch.empty_like(E_loc)
        m1 = torch.empty(2, dtype=torch.float32, device=dev)
        m2 = torch.empty(2, dtype=torch.float32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(lib.dpe_energy_moments1(p(E_loc), n, p(cw), clip_mode, p(E_clipped), p(m1), stream), "dpe_energy_moments1")
            m1 = utils.pmean(m1)          # E_mean, E_mean_clipped (loss_function.py:94, 21, 98)
            _lib.check(lib.dpe_energy_moments2(p(E_loc), p(E_clipped), n, p(m1), p(m2), stream), "dpe_energy_moments2")
            m2 = utils.pmean(m2)          # E_var, E_var_clipped (loss_function.py:95, 26-27, 99)
        E_mean, E_mean_clipped = m1[0], m1[1]
        E_var, E_var_clipped = m2[0], m2[1]
        new_state = (E_mean_clipped, torch.sqrt(E_var_clipped) * clipping_config.clip_by)   # loss_function.py:19-30
        aux = dict(E_mean=E_mean, E_var=E_var, E_mean_clipped=E_mean_clipped, E_var_clipped=E_var_clipped,
                   E_loc_clipped=E_clipped, E_loc=E_loc)
        return E_mean_clipped, (new_state, aux)

    return total_energy
