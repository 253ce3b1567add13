"""Checkpoint interop (deeperwin_b200/checkpoints.py vs the reference's checkpoints.py:25-93): reading a file written by the reference's
own save_run (tests/golden/reference_chkpt.zip, made by tests/golden/make_reference_checkpoint.py), round trips, pickled jax arrays,
device-split walker states, and that foreign classes are never executed.  The opposite direction -- the reference's load_run reading a
file written here -- is in tests/test_reference_pin.py."""
import io
import pickle
import sys
import types
import zipfile
from pathlib import Path

import numpy as np
import pytest
import torch

import deeperwin_b200 as dpe
from deeperwin_b200 import checkpoints as chk
from deeperwin_b200.mcmc import MCMCState

GOLDEN = Path(__file__).parent / "golden" / "reference_chkpt.zip"


def _state(B=8):
    g = torch.Generator().manual_seed(0)
    return MCMCState(r=torch.randn(B, 4, 3, generator=g), R=torch.randn(2, 3, generator=g), Z=torch.tensor([3, 1], dtype=torch.int32),
                     log_psi_sqr=torch.randn(B, generator=g), walker_age=torch.arange(B, dtype=torch.int32) % 3,
                     rng_state=torch.arange(2 * B, dtype=torch.int32).reshape(B, 2).view(torch.uint32),
                     stepsize=torch.tensor(0.3), step_nr=torch.tensor(17, dtype=torch.int32), acc_rate=torch.tensor(0.45))


def test_reads_a_checkpoint_written_by_the_reference():
    data = dpe.load_run(GOLDEN, device="cpu")
    assert data.config is None and data.metadata == dict(n_epochs=120, code_version="fixture") and data.fixed_params == {}
    st = data.mcmc_state
    assert isinstance(st, MCMCState) and st.r.shape == (24, 4, 3) and st.r.dtype == torch.float32
    assert st.rng_state.dtype == torch.uint32 and st.walker_age.dtype == torch.int32 and st.Z.tolist() == [3, 1]
    assert float(st.stepsize) == pytest.approx(0.37) and int(st.step_nr) == 120 and float(st.acc_rate) == pytest.approx(0.52)
    assert st.walker_age.tolist() == [i % 3 for i in range(24)]
    # the walkers are those of initialize_around_nuclei(24, LiH, seed 99): reproduce them with the oracle
    from oracle import mcmc as omc
    phys = dpe.PhysicalConfig(name="LiH")
    o = omc.initialize_around_nuclei(24, phys.R, phys.Z, phys.el_ion_mapping, 99, "gaussian", n_up=2)
    assert np.array_equal(st.rng_state.view(torch.int32).numpy().view(np.uint32), o.rng_state) and np.abs(st.r.numpy() - o.r).max() < 1e-6
    assert data.clipping_state == (pytest.approx(-8.01), pytest.approx(0.7))
    # parameters: the haiku tree of the small LiH model, numpy arrays
    from oracle import model as om
    d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3, n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)
    ref = om.cast_params(om.init_params(d, seed=31, bias_scale=0.1, envelope_jitter=0.5), torch.float32)
    assert set(data.params) == set(ref)
    for m, leaves in ref.items():
        for k, v in leaves.items():
            assert isinstance(data.params[m][k], np.ndarray) and np.array_equal(data.params[m][k], v.numpy())
    csv = dpe.load_run(GOLDEN, parse_csv=True, load_pkl=False)
    assert list(csv.history["opt_E_mean"]) == [-7.9, -8.0] and csv.params is None


def test_round_trip(tmp_path):
    cfg = dpe.Configuration(physical=dict(name="LiH"), optimization=dict(mcmc=dict(proposal=dict(name="local"))))
    st = _state()
    data = dpe.RunData(config=cfg, history=[dict(opt_epoch=0, E=1.0), dict(opt_epoch=1, E=0.5, extra=2)], summary=dict(E_mean=-8.0), metadata=dict(n_epochs=2),
                       params={"wf/a": {"w": torch.randn(3, 2), "b": torch.zeros(2)}}, ema_params={"wf/a": {"w": torch.ones(3, 2)}}, mcmc_state=st,
                       clipping_state=(torch.tensor(0.5), 2.0))
    fn = tmp_path / "chkpt000002.zip"
    dpe.save_run(fn, data)
    assert set(zipfile.ZipFile(fn).namelist()) == {"config.yml", "history.csv", "summary.csv", "metadata.pkl", "params.pkl", "ema_params.pkl", "mcmc_state.pkl",
                                                   "clipping_state.pkl"}
    assert "deeperwin" not in sys.modules and "deeperwin.mcmc" not in sys.modules          # the class-path stand-in is gone again
    back = dpe.load_run(fn, device="cpu")
    assert back.config == cfg and back.metadata == dict(n_epochs=2) and back.clipping_state == (0.5, 2.0)
    assert np.array_equal(back.params["wf/a"]["w"], data.params["wf/a"]["w"].numpy()) and np.array_equal(back.ema_params["wf/a"]["w"], np.ones((3, 2), np.float32))
    for k in chk._STATE_FIELDS:
        a, b = getattr(back.mcmc_state, k), getattr(st, k)
        assert a.dtype == b.dtype and torch.equal(a.view(torch.int32) if a.dtype == torch.uint32 else a, b.view(torch.int32) if b.dtype == torch.uint32 else b), k
    assert dpe.load_run(fn, parse_config=False, load_pkl=False).config["physical"]["name"] == "LiH"
    t = chk.params_to_torch(back.params, "cpu")
    assert torch.equal(t["wf/a"]["w"], data.params["wf/a"]["w"])


def _with_fake_modules(build):
    """Pickle `build(mods)` while fake third-party modules are registered, as a file written in the reference's environment would be."""
    names = ["jax", "jax._src", "jax._src.array", "kfac_jax", "kfac_jax._src", "kfac_jax._src.optimizer", "deeperwin", "deeperwin.mcmc"]
    mods = {n: types.ModuleType(n) for n in names}
    sys.modules.update(mods)
    try:
        return pickle.dumps(build(mods), protocol=4)
    finally:
        for n in names:
            sys.modules.pop(n, None)


def test_jax_arrays_device_split_states_and_foreign_classes(tmp_path):
    class _FakeJaxArray:                      # reduces exactly as jax._src.array.ArrayImpl.__reduce__ does
        def __init__(self, a):
            self.a = np.asarray(a)

        def __reduce__(self):
            fun, args, arr_state = self.a.__reduce__()
            return sys.modules["jax._src.array"]._reconstruct_array, (fun, args, arr_state, {"weak_type": False, "named_shape": {}})

    executed = []

    def build(mods):
        def _reconstruct_array(*a):
            raise AssertionError("not called while pickling")
        _reconstruct_array.__module__, _reconstruct_array.__qualname__ = "jax._src.array", "_reconstruct_array"
        mods["jax._src.array"]._reconstruct_array = _reconstruct_array

        class State:                          # a kfac_jax-like record whose constructor must never run on load
            def __init__(self):
                self.damping = 1e-3

            def __setstate__(self, s):
                executed.append(s)
        State.__module__, State.__qualname__ = "kfac_jax._src.optimizer", "State"
        mods["kfac_jax._src.optimizer"].State = State

        class MCMCStateRef:
            pass
        MCMCStateRef.__module__, MCMCStateRef.__qualname__ = "deeperwin.mcmc", "MCMCState"
        mods["deeperwin.mcmc"].MCMCState = MCMCStateRef
        st, D = _state(8), 2
        ref = MCMCStateRef()
        split = lambda t: _FakeJaxArray(chk._to_numpy(t).reshape((D, -1) + tuple(t.shape[1:])))
        tile = lambda t: _FakeJaxArray(np.stack([chk._to_numpy(t)] * D))
        ref.__dict__.update(r=split(st.r), log_psi_sqr=split(st.log_psi_sqr), walker_age=split(st.walker_age), rng_state=split(st.rng_state),
                            R=tile(st.R), Z=tile(st.Z), stepsize=tile(st.stepsize), step_nr=tile(st.step_nr), acc_rate=tile(st.acc_rate))
        return dict(mcmc_state=ref, opt_state=State(), params={"m": {"w": _FakeJaxArray(np.arange(6, dtype=np.float32).reshape(2, 3))}})

    blob = _with_fake_modules(build)
    out = chk._Unpickler(io.BytesIO(blob)).load()
    assert executed == []                                                     # nothing of the foreign class ran
    assert isinstance(out["opt_state"], chk.Opaque) and out["opt_state"]._path == ("kfac_jax._src.optimizer", "State") and out["opt_state"].state == {"damping": 1e-3}
    assert isinstance(out["params"]["m"]["w"], np.ndarray) and out["params"]["m"]["w"].tolist() == [[0, 1, 2], [3, 4, 5]]
    merged = chk.mcmc_state_from_reference(out["mcmc_state"], "cpu")          # split over two devices -> merged (mcmc.py:131-147)
    st = _state(8)
    assert torch.equal(merged.r, st.r) and torch.equal(merged.walker_age, st.walker_age) and torch.equal(merged.R, st.R)
    assert merged.stepsize.shape == () and float(merged.stepsize) == pytest.approx(0.3) and int(merged.step_nr) == 17
    with pytest.raises(NotImplementedError):                                  # opaque records are not silently re-written
        dpe.save_run(tmp_path / "x.zip", dpe.RunData(opt_state=out["opt_state"]))
    # arbitrary callables in a pickle are not resolved: os.system would come back as an inert Opaque class
    evil = pickle.dumps(__import__("os").system)
    assert issubclass(chk._Unpickler(io.BytesIO(evil)).load(), chk.Opaque)
