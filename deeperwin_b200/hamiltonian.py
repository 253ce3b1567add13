"""Drop-in for src/deeperwin/hamiltonian.py:272-291 (build_local_energy): the returned
`get_local_energy(trainable_params, spin_state, r, R, Z, fixed_params) -> E_loc[B]` evaluates
E_loc = -1/2 (1/2 lap L + 1/4 |grad L|^2) + V, L = log psi^2, with one forward-Laplacian pass on the GPU.
"""
from __future__ import annotations


def build_local_energy(log_psi_squared, is_complex=False, is_periodic=False, include_heg_background=False,
                       forward_lap=False, max_batch_size=64):
    """`forward_lap` / `max_batch_size` are accepted for signature parity: both branches of the reference
    (hamiltonian.py:206-216 folx forward-Laplacian, :234-267 jvp loop) define the same quantity, and the
    64-walker chunking of folx.batched_vmap is a memory workaround that the CUDA path replaces by
    workspace-sized chunks."""
    if is_complex or is_periodic or include_heg_background:
        raise NotImplementedError("complex / periodic local energies are outside the hot-path scope")
    engine = getattr(log_psi_squared, "engine", None)
    if engine is None:
        raise TypeError("build_local_energy needs the log_psi_sqr callable returned by deeperwin_b200.build_log_psi_squared")

    def get_local_energy(trainable_params, spin_state, r, R, Z, fixed_params=None, with_aux=False):
        n_up, n_dn = int(spin_state[0]), int(spin_state[1])
        if (n_up, n_dn) != (engine.n_up, engine.n_el - engine.n_up):
            raise ValueError(f"model was built for n_up={engine.n_up}, n_dn={engine.n_el - engine.n_up}")
        engine.set_params(trainable_params)
        engine.set_geometry(R, Z)
        engine.set_tao_cache(((fixed_params or {}).get("cache") or {}).get("taos"))
        return engine.local_energy(r, with_aux=with_aux)

    get_local_energy.engine = engine
    return get_local_energy
