"""BASELINE.json configs[2] in miniature: one set of weights shared by several H-chain geometries (weight sharing,
variational_optimization.py:210-442 picks ONE geometry per step): the same handle must switch geometry per call and
reproduce the oracle for each; and the segmented / packed GEMM paths of the forward pass agree with the SIMT path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_shared_weights_over_hchain_geometries():
    import deeperwin_b200 as dpe
    from oracle import model as om
    n_at = 6
    phys = [dpe.PhysicalConfig(name=f"HChain{n_at}_{a:.2f}", R=[[a * k, 0.0, 0.0] for k in range(n_at)], Z=[1] * n_at,
                               n_electrons=n_at, n_up=n_at // 2, el_ion_mapping=[0, 2, 4, 1, 3, 5]) for a in np.linspace(1.6, 2.6, 4)]
    cfg = dpe.Configuration(physical=phys[0].model_dump())
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys[0], None, None, rng_seed=3, device="cuda:0")
    gle = dpe.build_local_energy(f, forward_lap=True)
    d = om.ModelDims(n_el=n_at, n_up=n_at // 2, n_ion=n_at, Z_max=1)
    p64 = {m: {k: v.double().cpu() for k, v in l.items()} for m, l in params.items()}
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=5, initialization="gaussian"))
    states = [dpe.MCMCState.initialize_around_nuclei(48, p, "gaussian", "el_ion_mapping", dpe.PRNGKey(10 + g), device="cuda:0")
              for g, p in enumerate(phys)]
    for epoch in range(2):                                   # round-robin over geometries, as the shared optimisation does
        for g, p in enumerate(phys):
            states[g] = mc.run_inter_steps(f, states[g], params, p.n_up, p.n_dn, fixed)
            st = states[g]
            e = gle(params, (p.n_up, p.n_dn), st.r, st.R, st.Z, fixed).double().cpu()
            ref = om.forward_laplacian(p64, d, st.r.double().cpu(), st.R.double().cpu(), p.Z)
            lp = f(params, p.n_up, p.n_dn, st.r, st.R, st.Z, fixed)[1].double().cpu()
            assert ((lp - ref["logpsi2"]).abs() / ref["logpsi2"].abs()).median() < 1e-5
            rel = (e - ref["E_loc"]).abs() / ref["E_loc"].abs().clamp_min(1.0)
            assert rel.median() < 1e-4, (g, rel.median())
            assert torch.equal(lp.float(), st.log_psi_sqr.cpu())          # state carries the log psi^2 of ITS geometry
    assert int(states[0].step_nr) == 10


@pytest.mark.parametrize("name", ["LiH", "N2", "N"])
def test_tensor_core_and_simt_paths_agree(name):
    """gemm_path 1 (tcgen05 3xTF32) vs gemm_path 0 (FP32 SIMT) on the same walkers.  The tensor-core accumulator rounds
    towards zero, which makes a 3xTF32 dense layer 2-4x noisier than an FP32 FMA chain (2e-6 vs 7e-7 of max|C|); through the
    ill-conditioned determinants that is a few 1e-5 in E_loc on the median walker -- inside the 1e-4 budget, asserted here."""
    import deeperwin_b200 as dpe
    cfg = dpe.Configuration(physical=dict(name=name))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=5, device="cuda:0")
    st = dpe.MCMCState.initialize_around_nuclei(256, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1), device="cuda:0")
    st = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=40, initialization="gaussian")).run_inter_steps(
        f, st, params, phys.n_up, phys.n_dn, fixed)
    eng = f.engine
    if eng.lib.dpe_get_gemm_path(eng.handle) != 1:
        pytest.skip("tensor-core path unavailable")
    lp1 = f(params, phys.n_up, phys.n_dn, st.r, st.R, st.Z, fixed)[1]      # sets the parameters and the geometry
    e1 = eng.local_energy(st.r)
    eng.set_gemm_path(0)
    lp0 = eng.log_psi_sqr(st.r)[1]
    e0 = eng.local_energy(st.r)
    eng.set_gemm_path(1)
    assert ((lp1 - lp0).abs() / lp0.abs()).median() < 5e-6
    assert ((e1 - e0).abs() / e0.abs().clamp_min(1.0)).median() < 3e-4


def test_shared_optimization_loop_over_geometries():
    """shared_optimization_step (variational_optimization.py:354-400): scheduler, Metropolis inter-steps of the chosen geometry, loss + gradient
    on its walkers, the caller's update, EMA of the parameters, per-geometry bookkeeping."""
    import deeperwin_b200 as dpe
    n_at = 6
    phys = [dpe.PhysicalConfig(name=f"HChain{n_at}_{a:.2f}", R=[[a * k, 0.0, 0.0] for k in range(n_at)], Z=[1] * n_at,
                               n_electrons=n_at, n_up=n_at // 2, el_ion_mapping=[0, 2, 4, 1, 3, 5]) for a in np.linspace(1.6, 2.6, 3)]
    cfg = dpe.Configuration(physical=phys[0].model_dump())
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys[0], None, None, rng_seed=3, device="cuda:0")
    gle = dpe.build_local_energy(f, forward_lap=True)
    vag = dpe.build_value_and_grad_func(f, gle, dpe.ClippingConfig())
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=3, initialization="gaussian"))
    geoms = [dpe.GeometryDataStore(idx=g, physical_config=p, spin_state=(3, 3), fixed_params=fixed, clipping_state=dpe.init_clipping_state(),
                                   mcmc_state=dpe.MCMCState.initialize_around_nuclei(32, p, "gaussian", "el_ion_mapping", dpe.PRNGKey(g), device="cuda:0"))
             for g, p in enumerate(phys)]
    ema = {m: {k: v.clone() for k, v in l.items()} for m, l in params.items()}
    p0 = {m: {k: v.clone() for k, v in l.items()} for m, l in params.items()}
    seen_grads = []

    def update(p, grads, aux):
        seen_grads.append(grads)
        return {m: {k: v - 1e-3 * grads[m][k] for k, v in l.items()} for m, l in p.items()}

    order = []
    for epoch in range(9):
        method = "round_robin" if epoch < 6 else "stddev"
        params, idx, loss = dpe.shared_optimization_step(epoch, geoms, f, vag, mc, params, update, scheduling_method=method, n_initial_round_robin_per_geom=2,
                                                         permutation=[2, 0, 1], ema_params=ema, params_ema_factor=0.9)
        order.append(idx)
        assert torch.isfinite(loss)
    assert order[:6] == [2, 0, 1, 2, 0, 1]
    stds = [float(torch.sqrt(g.current_metrics["E_var"])) for g in geoms]
    assert all(g.n_opt_epochs >= 2 for g in geoms) and sum(g.n_opt_epochs for g in geoms) == 9
    assert int(geoms[order[-1]].mcmc_state.step_nr) == 3 * geoms[order[-1]].n_opt_epochs and geoms[order[-1]].last_epoch_optimized == 8
    # EMA: 0.9 ema + 0.1 params after every step, started from the initial parameters
    k = ("wf/~/input/h_ion", "embeddings")
    e = p0[k[0]][k[1]].clone()
    cur = p0[k[0]][k[1]].clone()
    for gr in seen_grads:
        cur = cur - 1e-3 * gr[k[0]][k[1]]
        e = 0.9 * e + 0.1 * cur
    assert torch.allclose(ema[k[0]][k[1]], e, rtol=1e-5, atol=1e-7)
    assert len(stds) == 3 and all(np.isfinite(stds))
