"""Weight sharing (BASELINE.json configs[2]): one set of weights optimised over several geometries, ONE geometry per step.

Host-side mirror of the pieces of the reference's shared loop that sit between the hot-path calls
(optimization/variational_optimization.py:300-400): the geometry record (geometries.py:15-38), the scheduler
`get_next_geometry_index` (utils/utils.py:704-748: round robin, then max-age / stddev / weight / var_per_el), the EMA of the parameters
(:389-393) and one step of the loop with the optimiser's update left to the caller.  Switching geometry costs two tiny device copies
(dpe_model_set_geometry_dev): no host round trip."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from .configuration import PhysicalConfig
from .mcmc import MCMCState, MetropolisHastingsMonteCarlo


@dataclass
class GeometryDataStore:
    """geometries.py:15-38 (the logger and distortion bookkeeping are control plane and not carried)."""
    idx: int = None
    physical_config: PhysicalConfig = None
    spin_state: Tuple[int, int] = None
    mcmc_state: MCMCState = None
    fixed_params: Dict = None
    clipping_state: Tuple[float, float] = None
    current_metrics: Dict[str, float] = dataclasses.field(default_factory=dict)
    n_opt_epochs: int = 0
    last_epoch_optimized: int = 0
    weight: float = None


def get_next_geometry_index(n_epoch: int, geometry_data_stores: List[GeometryDataStore], scheduling_method: str, max_age: Optional[int],
                            n_initial_round_robin_per_geom: int, permutation: Optional[List[int]] = None, rng: Optional[np.random.Generator] = None) -> int:
    """utils/utils.py:704-748.  (1) round robin -- always, or for the first n_initial_round_robin_per_geom rounds; (2) any geometry older
    than max_age (default 4 x number of geometries) goes first; (3) "stddev": largest sqrt(E_var); "weight" / "var_per_el": random with those
    probabilities.  Without a permutation the round-robin index itself is returned (the reference always passes one)."""
    n = len(geometry_data_stores)
    if scheduling_method == "round_robin" or n_epoch < n * n_initial_round_robin_per_geom:
        idx_next = n_epoch % n
        return int(permutation[idx_next]) if permutation is not None else idx_next
    ages = n_epoch - np.array([g.last_epoch_optimized for g in geometry_data_stores])
    max_age = max_age or int(n * 4.0)
    if np.any(ages > max_age):
        return int(np.argmax(ages))
    if scheduling_method == "stddev":
        return int(np.argmax([np.sqrt(float(g.current_metrics["E_var"])) for g in geometry_data_stores]))      # reads device scalars: synchronises
    rng = rng or np.random
    if scheduling_method == "weight":
        return int(rng.choice(n, p=[g.weight for g in geometry_data_stores]))
    if scheduling_method == "var_per_el":
        v = np.array([float(g.current_metrics.get("E_var", 1.0)) / g.physical_config.n_electrons for g in geometry_data_stores])
        return int(rng.choice(n, p=v / np.sum(v)))
    raise NotImplementedError("Wavefunction scheduler currently not supported.")


def update_ema_params(ema_params, params, factor: float):
    """variational_optimization.py:389-393: ema = factor * ema + (1 - factor) * params, in place, one multi-tensor kernel."""
    old = [t for leaves in ema_params.values() for t in leaves.values()]
    new = [params[m][k] for m, leaves in ema_params.items() for k in leaves]
    torch._foreach_lerp_(old, new, 1.0 - factor)
    return ema_params


def shared_optimization_step(n_epoch: int, geometries: List[GeometryDataStore], log_psi_sqr: Callable, value_and_grad: Callable,
                             mcmc: MetropolisHastingsMonteCarlo, params, optimizer_update: Callable[[Any, Any, Dict], Any], *,
                             scheduling_method="round_robin", max_age=None, n_initial_round_robin_per_geom=10, permutation=None,
                             ema_params=None, params_ema_factor=0.95, rng=None):
    """One pass through variational_optimization.py:354-400: pick the geometry, run its Metropolis inter-steps, evaluate loss + gradient
    (+ KFAC statistics) on its walkers, hand them to `optimizer_update(params, grads, aux) -> params`, update the EMA and the geometry's
    metrics.  Returns (params, index of the geometry, loss)."""
    idx = get_next_geometry_index(n_epoch, geometries, scheduling_method, max_age, n_initial_round_robin_per_geom, permutation, rng)
    g = geometries[idx]
    n_up, n_dn = g.spin_state
    g.mcmc_state = mcmc.run_inter_steps(log_psi_sqr, g.mcmc_state, params, n_up, n_dn, g.fixed_params)
    (loss, (g.clipping_state, aux)), grads = value_and_grad(params, g.clipping_state, g.spin_state, g.mcmc_state.build_batch(g.fixed_params))
    params = optimizer_update(params, grads, aux)
    if ema_params is not None:
        update_ema_params(ema_params, params, params_ema_factor)
    g.current_metrics = {k: v for k, v in aux.items() if not k.startswith("E_loc") and k != "kfac"}     # device scalars: read when needed
    g.n_opt_epochs += 1
    g.last_epoch_optimized = n_epoch
    return params, idx, loss
