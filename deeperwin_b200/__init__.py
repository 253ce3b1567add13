"""deeperwin_b200: B200-native (sm_100a) VMC inner loop behind DeepErwin's Python callables.

    log_psi_sqr, _, _, params, fixed = build_log_psi_squared(model_config, physical_config, None, None, seed)
    get_local_energy = build_local_energy(log_psi_sqr, forward_lap=True)
    mcmc = MetropolisHastingsMonteCarlo(mcmc_config); state = mcmc.run_inter_steps(log_psi_sqr, state, params, n_up, n_dn, fixed)

Importing the package does not load the CUDA library; the first call that needs it raises if it is missing
(there is no CPU fallback)."""
from .checkpoints import RunData, load_run, save_run
from .configuration import (ClippingConfig, Configuration, MCMCConfigEvaluation, MCMCConfigOptimization,
                            ModelConfigDeepErwin4, PhysicalConfig)
from .hamiltonian import build_local_energy
from .loss_function import build_total_energy, init_clipping_state
from .mcmc import MCMCState, MetropolisHastingsMonteCarlo, PRNGKey
from .optimization import build_value_and_grad_func
from .shared_optimization import GeometryDataStore, get_next_geometry_index, shared_optimization_step, update_ema_params
from .wavefunction import build_log_psi_squared

__all__ = ["Configuration", "PhysicalConfig", "ModelConfigDeepErwin4", "MCMCConfigOptimization", "MCMCConfigEvaluation",
           "ClippingConfig", "build_log_psi_squared", "build_local_energy", "build_total_energy", "init_clipping_state",
           "MCMCState", "MetropolisHastingsMonteCarlo", "PRNGKey", "build_value_and_grad_func", "RunData", "load_run", "save_run",
           "GeometryDataStore", "get_next_geometry_index", "shared_optimization_step", "update_ema_params"]
