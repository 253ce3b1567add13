"""Size-independent properties at BASELINE.json's full sizes (N2, 4096 walkers) and edge cases."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def n2():
    import deeperwin_b200 as dpe
    cfg = dpe.Configuration(physical=dict(name="N2"))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1234, device="cuda:0")
    state = dpe.MCMCState.initialize_around_nuclei(4096, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1234), device="cuda:0")
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=30, initialization="gaussian"))
    state = mc.run_inter_steps(f, state, params, phys.n_up, phys.n_dn, fixed)
    return dpe, phys, f, params, fixed, state


def test_full_size_antisymmetry_and_determinism(n2):
    dpe, phys, f, params, fixed, state = n2
    gle = dpe.build_local_energy(f, forward_lap=True)
    spin = (phys.n_up, phys.n_dn)
    r = state.r
    e1, a1 = gle(params, spin, r, state.R, state.Z, fixed, with_aux=True)
    e1b = gle(params, spin, r, state.R, state.Z, fixed)
    assert torch.equal(e1, e1b)                                              # deterministic: no atomics on the value path
    perm = list(range(phys.n_electrons)); perm[1], perm[4] = perm[4], perm[1]; perm[8], perm[12] = perm[12], perm[8]
    e2, a2 = gle(params, spin, r[:, perm], state.R, state.Z, fixed, with_aux=True)
    lp1, lp2 = a1["log_psi_sqr"], a2["log_psi_sqr"]
    rel_lp = (lp1 - lp2).abs() / lp1.abs()
    assert rel_lp.median() < 1e-5 and (rel_lp < 1e-5).float().mean() > 0.99 and rel_lp.max() < 1e-3, (rel_lp.median(), rel_lp.max())
    rel = (e1 - e2).abs() / e1.abs().clamp_min(1.0)
    assert rel.median() < 1e-4 and (rel < 1e-3).float().mean() > 0.9, (rel.median(), (rel < 1e-3).float().mean())
    g1 = a1["grad"].reshape(-1, phys.n_electrons, 3)
    g2 = a2["grad"].reshape(-1, phys.n_electrons, 3)[:, perm]                # gradient permutes with the electrons
    assert ((g1 - g2).abs().amax((1, 2)) / g1.abs().amax((1, 2))).median() < 1e-4
    ph1 = f(params, *spin, r, state.R, state.Z, fixed)[0]
    ph2 = f(params, *spin, r[:, [1, 0] + list(range(2, 14))], state.R, state.Z, fixed)[0]
    assert torch.all((ph1 - ph2).abs() > 3.0)                                # a single same-spin swap flips the sign
    assert torch.isfinite(e1).all() and e1.mean().item() < 0


def test_chunked_equals_single_pass(n2):
    """A workspace too small for the batch makes the library process ragged chunks: results are bitwise equal."""
    dpe, phys, f, params, fixed, state = n2
    eng = f.engine
    r = state.r[:1000]
    e_full, a_full = eng.local_energy(r, with_aux=True)
    lp_full = eng.log_psi_sqr(r)[1]
    cap, ws = eng.workspace_cap, eng._ws
    try:
        need = eng.lib.dpe_workspace_bytes(eng.handle, 1000, 1)
        eng.workspace_cap = int(need * 0.137)                                # 137-walker chunks + a ragged tail
        eng._ws = None
        e_ch, a_ch = eng.local_energy(r, with_aux=True)
        lp_ch = eng.log_psi_sqr(r)[1]
    finally:
        eng.workspace_cap, eng._ws = cap, ws
    assert torch.equal(e_full, e_ch) and torch.equal(a_full["grad"], a_ch["grad"]) and torch.equal(lp_full, lp_ch)


def test_mcmc_invariants_full_size(n2):
    dpe, phys, f, params, fixed, state = n2
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=3, initialization="gaussian"))
    new = mc.run_inter_steps(f, state, params, phys.n_up, phys.n_dn, fixed)
    age = new.walker_age
    assert int(age.min()) >= 0 and int(age.max()) <= 20
    counts = mc.last_accept_counts.cpu().numpy()
    assert counts.shape == (3,) and (counts > 0).all() and (counts <= 4096).all()
    moved = (new.r != state.r).any(-1).any(-1)
    assert int(moved.sum()) >= int(counts.max()) * 0 and int((~moved).sum()) == int((new.walker_age >= 3).sum())
    # a rejected-3-times walker kept its position and its log psi^2 equals a fresh forward pass
    lp = f(params, phys.n_up, phys.n_dn, new.r, new.R, new.Z, fixed)[1]
    assert torch.equal(lp, new.log_psi_sqr)
    assert int(new.step_nr) == int(state.step_nr) + 3
    # same input state -> same output state (per-walker counter RNG, no hidden state)
    again = mc.run_inter_steps(f, state, params, phys.n_up, phys.n_dn, fixed)
    assert torch.equal(again.r, new.r) and torch.equal(again.rng_state.view(torch.int32), new.rng_state.view(torch.int32))


def test_edge_cases():
    import deeperwin_b200 as dpe
    cfg = dpe.Configuration(physical=dict(name="LiH"))
    phys = cfg.physical
    f, get_slater, get_cache, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=0, device="cuda:0")
    st = dpe.MCMCState.initialize_around_nuclei(5, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(100), device="cuda:0")
    phase, lp = f(params, phys.n_up, phys.n_dn, st.r, st.R, st.Z, fixed)      # tests/test_forward_pass.py:13-49 of the reference
    assert lp is not None and lp.shape == (5,) and phase.shape == (5,)
    # single walker, extra batch dims (batch dims broadcast, input_features.py:121-125)
    p1, l1 = f(params, phys.n_up, phys.n_dn, st.r[0], st.R, st.Z, fixed)
    assert l1.shape == () and torch.equal(l1, lp[0])
    p2, l2 = f(params, phys.n_up, phys.n_dn, st.r.reshape(1, 5, 4, 3), st.R, st.Z, fixed)
    assert l2.shape == (1, 5) and torch.equal(l2[0], lp)
    gle = dpe.build_local_energy(f, forward_lap=True, max_batch_size=64)
    assert gle(params, (2, 2), st.r[:1], st.R, st.Z, fixed).shape == (1,)
    with pytest.raises(ValueError):
        f(params, 3, 1, st.r, st.R, st.Z, fixed)
    with pytest.raises(ValueError):
        f(params, 2, 2, st.r[:, :3], st.R, st.Z, fixed)
    with pytest.raises(NotImplementedError):
        dpe.build_local_energy(f, is_periodic=True)
    with pytest.raises(NotImplementedError):
        get_slater()
    assert get_cache() == {}
    # two electrons on top of each other / on a nucleus: non-finite energies surface as inf/nan, never as an error
    r_bad = st.r.clone(); r_bad[0, 1] = r_bad[0, 0]; r_bad[1, 0] = st.R[0]
    e = gle(params, (2, 2), r_bad, st.R, st.Z, fixed)
    assert not torch.isfinite(e[:2]).all() and torch.isfinite(e[2:]).all()
    # parameters are picked up when the caller changes them in place
    lp_a = f(params, 2, 2, st.r, st.R, st.Z, fixed)[1].clone()
    params["wf/~/orbitals/envelope_orbitals"]["alpha_up"].mul_(1.1)
    lp_b = f(params, 2, 2, st.r, st.R, st.Z, fixed)[1]
    assert not torch.equal(lp_a, lp_b)
    # a different geometry is picked up as well
    lp_c = f(params, 2, 2, st.r, st.R * 1.1, st.Z, fixed)[1]
    assert not torch.equal(lp_b, lp_c)
    n_walkers_zero = dpe.MCMCConfigOptimization(n_inter_steps=0, initialization="gaussian")
    s0 = dpe.MetropolisHastingsMonteCarlo(n_walkers_zero).run_inter_steps(f, st, params, 2, 2, fixed)
    assert torch.equal(s0.r, st.r) and int(s0.step_nr) == 0 and torch.equal(s0.log_psi_sqr, f(params, 2, 2, st.r, st.R, st.Z, fixed)[1])
