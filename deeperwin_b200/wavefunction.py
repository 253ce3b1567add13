"""Drop-in for the model callables of DeepErwin's wavefunction module.

Reference: src/deeperwin/model/wavefunction.py:262-299 (build_log_psi_squared) and :293 (the
`log_psi_sqr(params, n_up, n_dn, r, R, Z, fixed_params) -> (phase, log_psi_sqr)` callable that
mcmc.py:391, hamiltonian.py:208 and loss_function.py:144 consume).  Arrays are torch CUDA tensors instead
of jax arrays; all arithmetic runs in libdpe_b200.so.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .configuration import ModelConfigDeepErwin4, PhysicalConfig
from .engine import EMB, ORB, Engine, canonical_leaves


def construct_wavefunction_definition(config: ModelConfigDeepErwin4, physical_config: PhysicalConfig):
    """model/wavefunction.py:302-326: (Z_max, Z_min) of the lookup embedding."""
    Z_max = config.Z_max if config.Z_max is not None else int(max(physical_config.Z))
    Z_min = config.Z_min if config.Z_min is not None else int(min(physical_config.Z))
    return Z_max, Z_min


def _truncated_normal(shape, std, g):
    """hk.initializers.TruncatedNormal(std): a unit normal truncated at +-2, scaled by std."""
    t = torch.randn(shape, generator=g)
    for _ in range(8):
        bad = t.abs() > 2
        if not bad.any():
            break
        t = torch.where(bad, torch.randn(shape, generator=g), t)
    return t.clamp(-2, 2) * std


def init_params(engine: Engine, rng_seed: int, device, mlp_config=None) -> Dict[str, Dict[str, torch.Tensor]]:
    """Random initial weights with the reference's distributions: hk.initializers.VarianceScaling(1.0, mlp.init_weights_scale,
    mlp.init_weights_distribution) weights and TruncatedNormal(mlp.init_bias_scale) biases (mlp.py:42-43; defaults fan_avg /
    uniform / 0), envelope_orbitals.py:34-37: alpha = weights = 1, hk.Embed: truncated normal.  The random stream is torch's,
    not haiku's."""
    scale_mode = getattr(mlp_config, "init_weights_scale", "fan_avg")
    dist = getattr(mlp_config, "init_weights_distribution", "uniform")
    bias_scale = float(getattr(mlp_config, "init_bias_scale", 0.0))
    g = torch.Generator().manual_seed(int(rng_seed))
    params: Dict[str, Dict[str, torch.Tensor]] = {}
    for (mod, name), (_, _, rows, cols) in zip(engine.leaves, engine.leaf_shapes):
        if name == "w":
            n = {"fan_in": rows, "fan_out": cols, "fan_avg": 0.5 * (rows + cols)}[scale_mode]
            if dist == "uniform":
                t = (torch.rand(rows, cols, generator=g) * 2 - 1) * math.sqrt(3.0 / n)
            elif dist == "normal":
                t = torch.randn(rows, cols, generator=g) * math.sqrt(1.0 / n)
            else:       # truncated_normal: haiku rescales by the standard deviation of the truncated unit normal
                t = _truncated_normal((rows, cols), math.sqrt(1.0 / n) / 0.87962566103423978, g)
        elif name == "b":
            t = _truncated_normal((cols,), bias_scale, g) if bias_scale else torch.zeros(cols)
        elif name == "embeddings":
            t = _truncated_normal((rows, cols), 1.0, g)
        else:
            t = torch.ones(rows, cols)
        params.setdefault(mod, {})[name] = t.to(device=device, dtype=torch.float32)
    return params


def build_log_psi_squared(config: ModelConfigDeepErwin4, physical_config: PhysicalConfig, baseline_config=None,
                          fixed_params: Optional[Dict] = None, rng_seed: int = 0, device=None, workspace_gb: float = 48.0):
    """Same return tuple as the reference: (log_psi_sqr, get_slater_mat, get_cache, params, fixed_params)."""
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    emb, orb = config.embedding, config.orbitals
    Z_max, Z_min = construct_wavefunction_definition(config, physical_config)
    engine = Engine(n_el=physical_config.n_electrons, n_up=physical_config.n_up, n_ion=len(physical_config.Z),
                    n_iterations=emb.n_iterations, n_hidden_one_el=list(emb.n_hidden_one_el)[:emb.n_iterations],
                    n_hidden_two_el=list(emb.n_hidden_two_el)[:emb.n_iterations - 1], emb_dim=emb.emb_dim,
                    n_ion_features=config.features.n_ion_features, n_dets=orb.n_determinants, z_min=Z_min, z_max=Z_max,
                    use_taos=orb.transferable_atomic_orbitals is not None, device=device, workspace_gb=workspace_gb)
    params = init_params(engine, rng_seed, device, config.mlp)
    fixed_params = fixed_params if fixed_params is not None else {}

    def log_psi_sqr(params, n_up, n_dn, r, R, Z, fixed_params=None):
        if (n_up, n_dn) != (engine.n_up, engine.n_el - engine.n_up):
            raise ValueError(f"model was built for n_up={engine.n_up}, n_dn={engine.n_el - engine.n_up}")
        engine.set_params(params)
        engine.set_geometry(R, Z)
        # TAO models read the geometry cache exactly where the reference does (orbital_net.py:84-95)
        engine.set_tao_cache(((fixed_params or {}).get("cache") or {}).get("taos"))
        return engine.log_psi_sqr(r)

    def get_slater_mat(*args, **kwargs):
        raise NotImplementedError("Slater-matrix output is only used by pre-training (out of the hot-path scope)")

    def get_cache(*args, **kwargs):
        if engine.use_taos:
            raise NotImplementedError("the geometry-only TAO nets (TAOBackflow / TAOExponents, wavefunction.py:164-209) are a 'next' row; "
                                      'pass the cache in fixed_params["cache"]["taos"]')
        return {}   # envelope orbitals have no geometry-only cache (wavefunction.py:165-211 caches TAOs only)

    log_psi_sqr.engine = engine
    return log_psi_sqr, get_slater_mat, get_cache, params, fixed_params
