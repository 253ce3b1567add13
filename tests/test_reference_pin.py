"""Pins oracle/ against OUTPUTS OF THE REFERENCE ITSELF: the unmodified modules under /root/reference/src/deeperwin are imported
through tests/ref_shim (torch-backed stand-ins for jax / haiku / chex / folx, float64 on the CPU) and executed as they are.
Every quantity of the hot path is compared: the haiku parameter tree, log psi^2 and the phase (model/wavefunction.py),
the potential and kinetic energy and E_loc (hamiltonian.py, the forward_lap=False branch -- folx is not installed, both
branches define the same quantity), the Metropolis step (mcmc.py) and the clipped energy statistics (loss_function.py),
the walker initialisation (mcmc.py / orbitals.py) and the el_ion_mapping default (configuration.py).

Skipped where /root/reference does not exist (the GPU box): there the committed fixtures tests/golden/reference_*.npz, written
by tests/golden/make_reference_golden.py from the same execution, carry the reference's outputs (tests/test_gpu_parity.py)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present on this machine")


@pytest.fixture(scope="module")
def ref():
    """The reference's modules, imported unmodified under the shim."""
    ref_shim.install()
    import haiku as hk
    import deeperwin.configuration as cfg
    import deeperwin.hamiltonian as ham
    import deeperwin.mcmc as mcmc
    import deeperwin.model.wavefunction as wf
    import deeperwin.optimization.loss_function as loss
    import deeperwin.orbitals as orbitals
    import deeperwin.utils.utils as utils
    from deeperwin.model.definitions import WavefunctionDefinition
    from types import SimpleNamespace
    return SimpleNamespace(hk=hk, cfg=cfg, ham=ham, mcmc=mcmc, wf=wf, loss=loss, orbitals=orbitals, utils=utils, WfDef=WavefunctionDefinition)


def build_reference_model(ref, Z_max, **model_kw):
    """hk.multi_transform of the reference's Wavefunction, exactly as build_log_psi_squared does (model/wavefunction.py:275-293)."""
    cfg = ref.cfg.ModelConfigDeepErwin4(**model_kw)
    model = ref.hk.multi_transform(lambda: ref.wf.Wavefunction(cfg, ref.WfDef(Z_max=Z_max, Z_min=1)).init_for_multitransform())
    log_psi_sqr = lambda params, n_up, n_dn, *batch: model.apply[0](params, None, n_up, n_dn, *batch)
    return cfg, model, log_psi_sqr


def system(name, B, seed=0, **dims_kw):
    import deeperwin_b200 as dpe
    from oracle import model as om
    phys = dpe.PhysicalConfig(name=name)
    d = om.ModelDims(n_el=phys.n_electrons, n_up=phys.n_up, n_ion=len(phys.Z), Z_max=max(phys.Z), **dims_kw)
    p = om.init_params(d, seed=3 + seed, bias_scale=0.1, envelope_jitter=0.5)
    R = torch.tensor(phys.R, dtype=torch.float64)
    g = torch.Generator().manual_seed(100 + seed)
    r = R[torch.tensor(phys.el_ion_mapping)][None] + torch.randn(B, d.n_el, 3, generator=g, dtype=torch.float64)
    return phys, d, p, R, torch.tensor(phys.Z), r


SMALL = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)


def small_model_kw():
    return dict(embedding=dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], n_hidden_el_ions=[4], emb_dim=8),
                orbitals=dict(n_determinants=3))


@pytest.mark.parametrize("name,small", [("LiH", False), ("N2", False), ("B", False), ("HChain6", False), ("LiH", True), ("Ethene", True)])
def test_parameter_tree_and_log_psi_sqr(ref, name, small):
    """The reference's own Wavefunction creates exactly the oracle's parameter tree (haiku names, shapes), and with the same weights
    its log psi^2 / phase equal the oracle's to round-off."""
    from oracle import model as om
    phys, d, p, R, Z, r = system(name, 6, **(SMALL if small else {}))
    _, model, log_psi_sqr = build_reference_model(ref, max(phys.Z), **(small_model_kw() if small else {}))
    created = model.init(None, phys.n_up, phys.n_dn, r[:1], R, Z, {})
    assert {m: {k: tuple(v.shape) for k, v in l.items()} for m, l in created.items()} == \
           {m: {k: tuple(s) for k, s in l.items()} for m, l in om.param_shapes(d).items()}
    phase_ref, lp_ref = log_psi_sqr(p, phys.n_up, phys.n_dn, r, R, Z, {})
    phase, lp = om.log_psi_sqr(p, d, r, R, phys.Z)
    assert torch.allclose(lp, lp_ref, rtol=1e-12, atol=1e-12), (lp - lp_ref).abs().max()
    assert torch.equal(phase > 1, phase_ref > 1) and float((phase - phase_ref).abs().max()) < 1e-12
    # leading batch dimensions broadcast (input_features.py:121-125)
    _, lp2 = log_psi_sqr(p, phys.n_up, phys.n_dn, r.reshape(2, 3, d.n_el, 3), R, Z, {})
    assert torch.allclose(lp2.reshape(-1), lp_ref, rtol=1e-12, atol=1e-12)


def test_leaf_functions(ref):
    """get_distance_matrix, get_el_ion_distance_matrix (utils/utils.py:262-305), get_potential_energy (hamiltonian.py:17-39),
    evaluate_sum_of_determinants (model/wavefunction.py:63-83), residual_update (model/mlp.py:13-16)."""
    from oracle import model as om
    from deeperwin.model.mlp import residual_update
    phys, d, p, R, Z, r = system("N2", 5)
    diff_ee, dist_ee, diff_eI, dist_eI = om.distances(r, R)
    dr, di = ref.utils.get_distance_matrix(r)
    assert torch.equal(dr, diff_ee) and torch.allclose(di, dist_ee, rtol=1e-15, atol=0)
    dr, di = ref.utils.get_el_ion_distance_matrix(r, R)
    assert torch.equal(dr, diff_eI) and torch.allclose(di, dist_eI, rtol=1e-15, atol=0)
    assert torch.allclose(ref.ham.get_potential_energy(r, R, Z.double()), om.potential_energy(r, R, phys.Z), rtol=1e-13)
    g = torch.Generator().manual_seed(1)
    A = torch.randn(4, 32, 7, 7, generator=g, dtype=torch.float64)
    ph_ref, lp_ref = ref.wf.evaluate_sum_of_determinants(A[..., :3, :], A[..., 3:, :])
    ph, lp = om.sum_of_determinants(A)
    assert torch.allclose(lp, lp_ref, rtol=1e-13) and torch.allclose(ph, ph_ref)
    x, y = torch.randn(3, 8, dtype=torch.float64), torch.randn(3, 8, dtype=torch.float64)
    assert torch.equal(residual_update(y, x), om._res(y, x)) and torch.equal(residual_update(y, x[:, :4]), om._res(y, x[:, :4]))


@pytest.mark.parametrize("name,small,B", [("LiH", False, 4), ("LiH", True, 6), ("B", True, 4), ("N2", True, 2)])
def test_local_energy(ref, name, small, B):
    """E_loc of the reference's build_local_energy (hamiltonian.py:272-291; kinetic energy :234-267: jax.grad + jax.linearize +
    fori_loop over the 3N coordinates) equals the oracle's explicit forward-Laplacian E_loc."""
    from oracle import model as om
    phys, d, p, R, Z, r = system(name, B, **(SMALL if small else {}))
    _, _, log_psi_sqr = build_reference_model(ref, max(phys.Z), **(small_model_kw() if small else {}))
    get_local_energy = ref.ham.build_local_energy(log_psi_sqr, forward_lap=False, max_batch_size=64)
    e_ref = get_local_energy(p, (phys.n_up, phys.n_dn), r, R, Z.double(), {})
    out = om.forward_laplacian(p, d, r, R, phys.Z)
    assert torch.allclose(out["E_loc"], e_ref, rtol=1e-9, atol=1e-9), (out["E_loc"] - e_ref).abs().max()


def _ref_mcmc_config(ref, **kw):
    return ref.cfg.MCMCConfigOptimization(**kw)


@pytest.mark.parametrize("proposal,n_steps", [("normal", 12), ("normal_one_el", 9), ("cauchy", 4), ("local", 8), ("local_one_el", 9), ("langevin", 8)])
def test_metropolis_chain(ref, proposal, n_steps):
    """MetropolisHastingsMonteCarlo._run_mcmc_steps of the reference (mcmc.py:345-406) against the oracle chain on the same
    log psi^2 function: keys, ages and step counters bit-exact, positions / step size / acceptance rate to float32 round-off
    (the shim evaluates in float64, the oracle in the reference's float32)."""
    from oracle import mcmc as omc, model as om, threefry
    phys, d, p, R, Z, _ = system("LiH", 1)
    _, _, log_psi_sqr = build_reference_model(ref, 3)
    B = 16
    st0 = omc.initialize_around_nuclei(B, phys.R, phys.Z, phys.el_ion_mapping, 1234, "gaussian", n_up=phys.n_up)
    func = lambda rr: om.log_psi_sqr(p, d, torch.from_numpy(rr).double(), R, phys.Z)[1].float().numpy()
    pkw = {"local": dict(r_min=0.15, r_max=0.9), "local_one_el": dict(r_min=0.15, r_max=0.9), "langevin": dict(langevin_scale=0.7, r_min=0.25, r_max=1.5)}.get(proposal, {})
    oracle_state = omc.run_mcmc_steps(func, st0, n_steps, max_age=3, stepsize_update_interval=5, proposal=proposal, proposal_kw=pkw)
    cfg = _ref_mcmc_config(ref, n_inter_steps=n_steps, max_age=3, stepsize_update_interval=5, proposal=dict(name=proposal, **pkw))
    mc = ref.mcmc.MetropolisHastingsMonteCarlo(cfg)
    state = ref.mcmc.MCMCState(r=torch.from_numpy(st0.r).double(), R=R, Z=Z, log_psi_sqr=torch.from_numpy(st0.log_psi_sqr).double(),
                               walker_age=torch.from_numpy(st0.walker_age).long(), rng_state=st0.rng_state.copy())
    new = mc._run_mcmc_steps(log_psi_sqr, state, p, phys.n_up, phys.n_dn, {}, n_steps)
    assert np.array_equal(np.asarray(new.rng_state), oracle_state.rng_state)
    assert int(new.step_nr) == oracle_state.step_nr == n_steps
    same = new.walker_age.numpy() == oracle_state.walker_age
    assert same.mean() >= (1.0 if proposal != "cauchy" else 0.9)
    assert np.abs(new.r.numpy() - oracle_state.r)[same].max() < (2e-6 if proposal != "cauchy" else 1e-3)
    assert abs(float(new.stepsize) - float(oracle_state.stepsize)) < 1e-7 and abs(float(new.acc_rate) - float(oracle_state.acc_rate)) < 1e-6
    # the walker initialisation the chain started from is the reference's as well (mcmc.py:39-91)
    ref_phys = ref.cfg.PhysicalConfig(name=None, R=phys.R, Z=phys.Z, n_electrons=phys.n_electrons, n_up=phys.n_up, el_ion_mapping=phys.el_ion_mapping)
    ref_init = ref.mcmc.MCMCState.initialize_around_nuclei(B, ref_phys, "gaussian", "el_ion_mapping", threefry.prng_key(1234))
    assert np.array_equal(np.asarray(ref_init.rng_state), st0.rng_state) and np.abs(ref_init.r.numpy() - st0.r).max() < 1e-6


@pytest.mark.parametrize("name", ["LiH", "N2", "Benzene"])
def test_exponential_walker_initialisation(ref, name):
    """initialization = "exponential" (the reference's default, orbitals.py:854-928, utils/utils.py:387-506)."""
    import deeperwin_b200 as dpe
    from oracle import mcmc as omc, threefry
    phys = dpe.PhysicalConfig(name=name)
    ref_phys = ref.cfg.PhysicalConfig(name=None, R=phys.R, Z=phys.Z, n_electrons=phys.n_electrons, n_up=phys.n_up, el_ion_mapping=phys.el_ion_mapping)
    a = ref.mcmc.MCMCState.initialize_around_nuclei(32, ref_phys, "exponential", "el_ion_mapping", threefry.prng_key(77))
    b = omc.initialize_around_nuclei(32, phys.R, phys.Z, phys.el_ion_mapping, 77, "exponential", n_up=phys.n_up)
    assert np.array_equal(np.asarray(a.rng_state), b.rng_state)
    assert np.allclose(a.r.numpy(), b.r, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("name,center,width_metric,from_prev", [("tanh", "mean", "std", True), ("hard", "median", "mae", True),
                                                                 ("tanh", "median", "mae", False), ("hard", "mean", "std", False)])
def test_clipped_energy_statistics(ref, name, center, width_metric, from_prev):
    """_clip_energies and the statistics of total_energy (loss_function.py:19-109)."""
    from oracle import mcmc as omc
    cc = ref.cfg.ClippingConfig(name=name, center=center, width_metric=width_metric, from_previous_step=from_prev, clip_by=3.0)
    g = torch.Generator().manual_seed(7)
    for n, n_nan in ((257, 0), (256, 0), (300, 3)):
        E = (-10 + 2 * torch.randn(n, generator=g, dtype=torch.float64)).float().double()
        E[5], E[11] = 400.0, -300.0
        if n_nan:
            E[torch.arange(n_nan) * 17 + 1] = float("nan")
        state = (torch.tensor(-9.5, dtype=torch.float64), torch.tensor(4.0, dtype=torch.float64))
        Ec_ref, new_state_ref = ref.loss._clip_energies(E, state, cc)
        _, new_state, aux = omc.energy_statistics(E.numpy().astype(np.float32), (np.float32(-9.5), np.float32(4.0)), name=name, clip_by=3.0,
                                                  center=center, width_metric=width_metric, from_previous_step=from_prev)
        ok = ~torch.isnan(E)
        assert np.allclose(aux["E_loc_clipped"][ok.numpy()], Ec_ref[ok].numpy(), rtol=3e-6, atol=3e-5)
        assert abs(float(new_state[0]) - float(new_state_ref[0])) < 3e-5 * max(1.0, abs(float(new_state_ref[0])))
        assert abs(float(new_state[1]) - float(new_state_ref[1])) < 3e-5 * float(new_state_ref[1])
        # the full total_energy through the reference's value_and_grad wrapper (forward value and aux)
        vag = ref.loss.build_value_and_grad_func(lambda *a: None, lambda params, spin, *batch: E, cc)
        (loss, (st, stats)) = vag.fun({}, state, (1, 1), (None, None, None, {}))
        for k in ("E_mean", "E_var", "E_mean_clipped", "E_var_clipped"):
            assert abs(float(stats[k]) - float(aux[k])) <= 5e-5 * max(1.0, abs(float(aux[k]))), k


def test_el_ion_mapping_default(ref):
    """PhysicalConfig._generate_el_ion_mapping (configuration.py:1571-1615) vs deeperwin_b200.configuration.generate_el_ion_mapping."""
    import yaml
    from deeperwin_b200.configuration import generate_el_ion_mapping
    mols = yaml.safe_load((ref_shim.REFERENCE_SRC / "deeperwin" / "molecules.yaml").read_text())
    n = 0
    for name, m in mols.items():
        if "R" not in m:
            continue
        Z = m["Z"]
        n_el = sum(Z) - m.get("charge", 0)
        n_up = (n_el + m.get("spin", 0) + 1) // 2
        assert ref.cfg.PhysicalConfig._generate_el_ion_mapping(m["R"], Z, n_el, n_up) == generate_el_ion_mapping(m["R"], Z, n_el, n_up), name
        n += 1
    assert n > 20


def test_config_defaults_match(ref):
    """The config mirror's defaults equal the reference's for every field the hot path reads."""
    import deeperwin_b200.configuration as mine
    pairs = [(ref.cfg.MCMCConfigOptimization(), mine.MCMCConfigOptimization()), (ref.cfg.MCMCConfigEvaluation(), mine.MCMCConfigEvaluation()),
             (ref.cfg.ClippingConfig(), mine.ClippingConfig()), (ref.cfg.EmbeddingConfigDeepErwin4(), mine.EmbeddingConfigDeepErwin4()),
             (ref.cfg.InputFeatureConfigDPE4(), mine.InputFeatureConfigDPE4()), (ref.cfg.MLPConfig(), mine.MLPConfig()),
             (ref.cfg.EnvelopeOrbitalsConfig(), mine.EnvelopeOrbitalsConfig())]
    for a, b in pairs:
        da, db = a.model_dump(), b.model_dump()
        for k, v in db.items():
            if k in da:
                assert da[k] == v, (type(b).__name__, k, da[k], v)


def test_parameter_gradient(ref):
    """The gradient of the reference's own total_energy (optimization/loss_function.py:89-154: jax.value_and_grad of the function
    with the custom jvp) equals the oracle's backward pass of sum_b (E_c - mean E_c)_b / B * log psi^2_b."""
    from oracle import gradient as og
    phys, d, p, R, Z, r = system("LiH", 6, **SMALL)
    _, _, log_psi_sqr = build_reference_model(ref, 3, **small_model_kw())
    E = torch.tensor([-7.9, -8.3, -8.1, -7.5, -9.0, -8.2], dtype=torch.float64)
    cc = ref.cfg.ClippingConfig(name="hard", center="mean", width_metric="std", from_previous_step=True, clip_by=5.0)
    vag = ref.loss.build_value_and_grad_func(log_psi_sqr, lambda params, spin, *batch: E, cc)
    state = (torch.tensor(-8.0, dtype=torch.float64), torch.tensor(100.0, dtype=torch.float64))       # wide window: nothing is clipped
    (loss, (st, stats)), grads = vag(p, state, (phys.n_up, phys.n_dn), (r, R, Z, {}))
    mine = og.loss_gradient_from_energies(p, d, r, R, phys.Z, E)
    n = 0
    for m, leaves in mine.items():
        for k, g in leaves.items():
            gr = grads[m][k]
            assert torch.allclose(g, gr, rtol=1e-9, atol=1e-12), (m, k, (g - gr).abs().max())
            n += 1
    assert n == sum(len(v) for v in p.values()) and abs(float(loss) - float(E.mean())) < 1e-12


def test_reference_load_run_reads_a_checkpoint_written_here(ref, tmp_path):
    """deeperwin_b200.save_run -> the reference's own load_run (checkpoints.py:69-93): mcmc_state.pkl comes back as the reference's
    MCMCState dataclass, params / clipping state / metadata unchanged.  (No config.yml: the reference parses it with ruamel.yaml, absent here.)"""
    import deeperwin.checkpoints as rchk
    import deeperwin_b200 as dpe
    from deeperwin_b200.mcmc import MCMCState
    g = torch.Generator().manual_seed(3)
    st = MCMCState(r=torch.randn(6, 4, 3, generator=g), R=torch.randn(2, 3, generator=g), Z=torch.tensor([3, 1], dtype=torch.int32),
                   log_psi_sqr=torch.randn(6, generator=g), walker_age=torch.arange(6, dtype=torch.int32),
                   rng_state=torch.arange(12, dtype=torch.int32).reshape(6, 2).view(torch.uint32), stepsize=torch.tensor(0.2),
                   step_nr=torch.tensor(40, dtype=torch.int32), acc_rate=torch.tensor(0.5))
    params = {"wf/~/input/h_ion": {"embeddings": torch.randn(3, 32, generator=g)}, "wf/x/linear_0": {"w": torch.randn(4, 5, generator=g), "b": torch.zeros(5)}}
    fn = tmp_path / "chkpt.zip"
    dpe.save_run(fn, dpe.RunData(params=params, mcmc_state=st, clipping_state=(torch.tensor(-8.0), 0.5), metadata=dict(n_epochs=40),
                                 history=[dict(opt_epoch=0, opt_E_mean=-7.5)], summary=dict(E_mean=-8.0)))
    data = rchk.load_run(str(fn))
    assert type(data.mcmc_state) is ref.mcmc.MCMCState
    assert np.array_equal(data.mcmc_state.r, st.r.numpy()) and data.mcmc_state.r.dtype == np.float32
    assert np.array_equal(data.mcmc_state.rng_state, st.rng_state.view(torch.int32).numpy().view(np.uint32)) and data.mcmc_state.rng_state.dtype == np.uint32
    assert np.array_equal(data.mcmc_state.walker_age, np.arange(6)) and float(data.mcmc_state.stepsize) == pytest.approx(0.2) and int(data.mcmc_state.step_nr) == 40
    assert data.mcmc_state.build_batch({})[1].shape == (2, 3)                            # the reference's own method on the loaded object
    assert set(data.params) == set(params) and np.array_equal(data.params["wf/x/linear_0"]["w"], params["wf/x/linear_0"]["w"].numpy())
    assert tuple(float(x) for x in data.clipping_state) == (-8.0, 0.5) and data.metadata == dict(n_epochs=40)


def test_geometry_scheduler_matches_the_reference(ref):
    from types import SimpleNamespace
    """get_next_geometry_index (utils/utils.py:704-748) of the reference vs deeperwin_b200.shared_optimization on the same records, for
    every deterministic branch: round robin (with the permutation the shared loop always passes), the max-age override, stddev."""
    import deeperwin_b200 as dpe
    rng = np.random.default_rng(0)
    n = 6
    for trial in range(40):
        stores_ref, stores = [], []
        for k in range(n):
            kw = dict(last_epoch_optimized=int(rng.integers(0, 60)), weight=1.0 / n)
            a, b = SimpleNamespace(**kw, current_metrics={"E_var": float(rng.uniform(0.1, 2.0))}), dpe.GeometryDataStore(**kw)
            b.current_metrics = dict(a.current_metrics)
            stores_ref.append(a); stores.append(b)
        perm = list(rng.permutation(n))
        n_epoch = int(rng.integers(0, 90))
        for method, max_age, n_rr in (("round_robin", None, 10), ("stddev", None, 2), ("stddev", 12, 0), ("stddev", 500, 0)):
            want = int(ref.utils.get_next_geometry_index(n_epoch, stores_ref, method, max_age, n_rr, perm))
            got = dpe.get_next_geometry_index(n_epoch, stores, method, max_age, n_rr, perm)
            assert got == want, (trial, method, max_age, n_rr, n_epoch)
