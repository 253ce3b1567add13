"""The optimisation-step pieces that belong to the hot path: the parameter gradient of the VMC loss and the KFAC curvature
statistics, with their cross-GPU reduction.

Reference: optimization/loss_function.py:75-154 (`build_value_and_grad_func`: total_energy with its custom jvp, wrapped in
jax.value_and_grad), optimizers.py:133 (pmean of the gradients) and kfac_jax's curvature update (custom_kfac_jax/kfac_jax/_src/
optimizer.py:1104-1189, curvature_blocks.py:1594-1624).  The KFAC inverses / preconditioning / parameter update are the optimiser's
control flow and stay outside."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist

from . import utils
from .configuration import ClippingConfig
from .loss_function import build_total_energy


def _unflatten(engine, flat, like) -> Dict[str, Dict[str, torch.Tensor]]:
    out: Dict[str, Dict[str, torch.Tensor]] = {}
    for (mod, name), (off, size, rows, cols) in zip(engine.leaves, engine.leaf_shapes):
        out.setdefault(mod, {})[name] = flat[off:off + size].reshape(like[mod][name].shape)
    return out


def build_value_and_grad_func(log_psi_sqr_func, get_local_energy, clipping_config: ClippingConfig, is_complex=False, kfac_register_complex=False,
                              with_kfac_statistics: bool = False):
    """Drop-in for loss_function.py:75-154: returns value_and_grad(params, state, spin_state, batch) -> ((loss, (state, aux)), grads)
    with `grads` in the haiku tree layout of `params`.  The gradient is  (1 / B) sum_b (E_clipped_b - mean E_clipped) d log psi^2_b,
    averaged over the GPUs.  With `with_kfac_statistics` the same backward pass also yields the Kronecker factors of every dense
    layer (aux["kfac"] = {haiku module: (A, G)}); gradient and factors travel through ONE flat NCCL all-reduce."""
    if is_complex or kfac_register_complex:
        raise NotImplementedError("complex wavefunctions are outside the hot-path scope")
    engine = getattr(log_psi_sqr_func, "engine", None)
    if engine is None:
        raise TypeError("build_value_and_grad_func needs the log_psi_sqr callable returned by deeperwin_b200.build_log_psi_squared")
    total_energy = build_total_energy(get_local_energy, clipping_config)

    def value_and_grad(params, state, spin_state, batch):
        loss, (new_state, aux) = total_energy(params, state, spin_state, batch)
        r, R, Z, fixed_params = batch
        diff = aux["E_loc_clipped"] - aux["E_mean_clipped"]                         # loss_function.py:145 (E_mean_clipped is the pmean)
        cot = torch.nan_to_num(diff, nan=0.0) / diff.numel()
        engine.set_params(params)
        engine.set_geometry(R, Z)
        engine.set_tao_cache(((fixed_params or {}).get("cache") or {}).get("taos"))      # TAO models: gradient of the embedding, cache held fixed
        flat, _ = engine.param_gradient(r, cot, with_kfac=with_kfac_statistics)
        if utils.world_size() > 1:                                                  # one flat all-reduce for gradient + KFAC factors
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat /= utils.world_size()
        grads = _unflatten(engine, flat, params)
        if with_kfac_statistics:
            kf = flat[engine.n_params:]
            aux["kfac"] = {name: (kf[a:a + (din + hb) ** 2].reshape(din + hb, din + hb), kf[g:g + dout * dout].reshape(dout, dout))
                           for name, din, dout, hb, _, a, g in engine.kfac_layers()}
        return (loss, (new_state, aux)), grads

    value_and_grad.engine = engine
    return value_and_grad
