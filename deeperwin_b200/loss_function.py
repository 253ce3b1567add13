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


def build_total_energy(get_local_energy, clipping_config: ClippingConfig, stats_stream: "torch.cuda.Stream | None" = None):
    """Returns total_energy(params, clipping_state, spin_state, batch) -> (loss, (clipping_state, aux)) with the
    aux keys of loss_function.py:100-107.  All clipping variants of the reference: tanh / hard window (:44-59), centre
    mean / median, width std / mae (:19-30), window from the previous step or from the current energies (:62-72).

    `stats_stream` (optional, multi-GPU): the reductions and their NCCL all-reduces are issued on that side stream, ordered after
    the E_loc pass by an event.  The scalar all-reduces are device-side rendezvous points of all ranks; on the compute stream each
    of them makes every GPU wait for the slowest one, on a side stream the next Metropolis steps run underneath.  The caller
    synchronises `stats_stream` (or waits on it) before reading the returned scalars."""
    clip_mode = {"tanh": 0, "hard": 1}[clipping_config.name]
    lib = _lib.load()
    p = lambda t: C.c_void_p(t.data_ptr())

    def center_and_width(E, n, stream, mean_and_var=None):
        """_get_clipping_center_and_width (loss_function.py:19-30); `mean_and_var` = already all-reduced (nanmean, variance
        about it) of E, reused when centre = mean and width = std."""
        dev = E.device
        if clipping_config.center == "mean":
            if mean_and_var is not None:
                center = mean_and_var[0].reshape(1)
            else:
                m = torch.empty(2, dtype=torch.float32, device=dev)
                cw0 = torch.tensor([0.0, 1e12], dtype=torch.float32, device=dev)
                _lib.check(lib.dpe_energy_moments1(p(E), n, p(cw0), 1, p(torch.empty_like(E)), p(m), stream), "dpe_energy_moments1")
                center = utils.pmean(m[:1])
        else:
            center = torch.empty(1, dtype=torch.float32, device=dev)
            _lib.check(lib.dpe_energy_median(p(E), n, p(center), stream), "dpe_energy_median")
            center = utils.pmean(center)                      # pmean of the per-device medians, as the reference
        if clipping_config.width_metric == "std" and clipping_config.center == "mean" and mean_and_var is not None:
            width = torch.sqrt(mean_and_var[1])
        else:
            w = torch.empty(1, dtype=torch.float32, device=dev)
            center = center.contiguous()
            _lib.check(lib.dpe_energy_width(p(E), n, p(center), int(clipping_config.width_metric == "mae"), p(w), stream), "dpe_energy_width")
            w = utils.pmean(w)
            width = w[0] if clipping_config.width_metric == "mae" else torch.sqrt(w[0])
        return center.reshape(()), width.reshape(()) * clipping_config.clip_by

    def total_energy(params, state, spin_state, batch):
        r, R, Z, fixed_params = batch
        E_loc = get_local_energy(params, spin_state, r, R, Z, fixed_params).reshape(-1).contiguous()
        if stats_stream is None:
            return _statistics(E_loc, state)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(E_loc.device))
        E_loc.record_stream(stats_stream)
        with torch.cuda.stream(stats_stream):
            stats_stream.wait_event(done)
            return _statistics(E_loc, state)

    def _statistics(E_loc, state):
        dev = E_loc.device
        n = E_loc.numel()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        E_clipped = torch.empty_like(E_loc)
        m1 = torch.empty(2, dtype=torch.float32, device=dev)
        m2 = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if clipping_config.from_previous_step and state is not None and state[0] is not None:
                cw = torch.stack([state[0].reshape(()), state[1].reshape(())]).to(torch.float32).contiguous()
            else:                                             # loss_function.py:65-66: window from the current energies
                cw = torch.stack(center_and_width(E_loc, n, stream)).to(torch.float32).contiguous()
            _lib.check(lib.dpe_energy_moments1(p(E_loc), n, p(cw), clip_mode, p(E_clipped), p(m1), stream), "dpe_energy_moments1")
            m1 = utils.pmean(m1)          # E_mean, E_mean_clipped (loss_function.py:94, 21, 98)
            _lib.check(lib.dpe_energy_moments2(p(E_loc), p(E_clipped), n, p(m1), p(m2), stream), "dpe_energy_moments2")
            m2 = utils.pmean(m2)          # E_var, E_var_clipped (loss_function.py:95, 26-27, 99)
            E_mean, E_mean_clipped = m1[0], m1[1]
            E_var, E_var_clipped = m2[0], m2[1]
            new_state = center_and_width(E_clipped, n, stream, mean_and_var=(E_mean_clipped, E_var_clipped))   # :70
        aux = dict(E_mean=E_mean, E_var=E_var, E_mean_clipped=E_mean_clipped, E_var_clipped=E_var_clipped,
                   E_loc_clipped=E_clipped, E_loc=E_loc)
        return E_mean_clipped, (new_state, aux)

    return total_energy
