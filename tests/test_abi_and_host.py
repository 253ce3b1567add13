"""CPU-side checks: the C-ABI library loads and exports every symbol include/dpe_b200.h declares (no compute
calls without a GPU), the config mirror behaves like the reference's pydantic tree, host-side scheduling logic."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "dpe_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dpe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(str(built_library))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libdpe_b200.so does not export {s}"


def test_ctypes_signatures_cover_header(built_library):
    from deeperwin_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert b"sm_100a" in lib.dpe_version()


def test_argument_errors_without_gpu(built_library):
    """Entry points validate arguments before touching the device, and report through dpe_last_error."""
    from deeperwin_b200 import _lib
    lib = _lib.load()
    d = _lib.DpeDims()
    h = ctypes.c_void_p()
    assert lib.dpe_model_create(None, ctypes.byref(h)) == -1
    d.n_el, d.n_up, d.n_ion, d.n_iterations = 4, 0, 2, 4       # no spin-up electron: mean over an empty slice
    assert lib.dpe_model_create(ctypes.byref(d), ctypes.byref(h)) == -2
    assert b"spin" in lib.dpe_last_error()
    assert lib.dpe_energy_moments1(None, 0, None, 0, None, None, None) == -1
    assert lib.dpe_param_count(None) == 0 and lib.dpe_workspace_bytes(None, 10, 0) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never routes through oracle/."""
    import deeperwin_b200 as dpe
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    cfg = dpe.Configuration(physical=dict(name="LiH"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dpe.build_log_psi_squared(cfg.model, cfg.physical, None, None, 0, device="cuda:0")
    for f in (ROOT / "deeperwin_b200").rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+\.*oracle", f.read_text(), re.M), f"{f} imports oracle/"


def test_configuration_mirror(tmp_path):
    import deeperwin_b200 as dpe
    cfg = dpe.Configuration(physical=dict(name="N2"))
    assert (cfg.physical.n_electrons, cfg.physical.n_up, cfg.physical.n_ions) == (14, 7, 2)
    assert cfg.physical.el_ion_mapping == [0, 0, 0, 0, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1]
    e = cfg.model.embedding
    assert (e.n_iterations, e.n_hidden_one_el, e.n_hidden_two_el, e.emb_dim) == (4, [256] * 4, [32] * 3, 32)
    assert cfg.model.orbitals.n_determinants == 32 and cfg.model.orbitals.determinant_schema == "full_det"
    m = cfg.optimization.mcmc
    assert (m.n_inter_steps, m.n_burn_in, m.max_age, m.stepsize_update_interval, m.n_walkers) == (20, 1000, 20, 100, 2048)
    assert cfg.evaluation.mcmc.max_age == 100 and cfg.optimization.clipping.clip_by == 5.0
    with pytest.raises(Exception):                        # extra="forbid" (configuration.py:91-95)
        dpe.Configuration(physical=dict(name="N2", foo=1))
    with pytest.raises(NotImplementedError):              # options the CUDA path does not implement raise loudly
        dpe.Configuration(model=dict(embedding=dict(use_h_two_same_diff=False)))
    with pytest.raises(Exception):
        dpe.Configuration(optimization=dict(mcmc=dict(proposal=dict(name="langevin"))))       # not among the simple proposals
    assert dpe.Configuration(optimization=dict(mcmc=dict(proposal=dict(name="cauchy")))).optimization.mcmc.proposal.name == "cauchy"
    p = tmp_path / "config.yml"
    cfg.save(p)
    assert dpe.Configuration.load_configuration_file(p).model_dump() == cfg.model_dump()
    he = dpe.PhysicalConfig(name="He")
    assert (he.Z, he.n_electrons, he.n_up, he.R) == ([2], 2, 1, [[0.0, 0.0, 0.0]])
    bz = dpe.PhysicalConfig(name="Benzene")
    assert (bz.n_electrons, bz.n_up, bz.n_ions) == (42, 21, 12)


def test_plan_segments():
    from deeperwin_b200.mcmc import plan_segments
    assert plan_segments(0, 20, 100) == [20]
    assert plan_segments(95, 20, 100) == [5, 15]
    assert plan_segments(100, 250, 100) == [100, 100, 50]
    assert plan_segments(7, 0, 100) == []
    for s0, n, iv in ((3, 57, 10), (0, 1000, 100), (99, 2, 100)):
        segs = plan_segments(s0, n, iv)
        assert sum(segs) == n
        pos = s0
        for seg in segs[:-1]:
            pos += seg
            assert pos % iv == 0        # every segment but the last ends on a step-size update


def test_canonical_leaf_order_matches_oracle_tree():
    from deeperwin_b200.engine import canonical_leaves
    from oracle import model as om
    d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3)
    shapes = om.param_shapes(d)
    leaves = canonical_leaves(d.n_iterations)
    assert len(leaves) == sum(len(v) for v in shapes.values())
    for mod, name in leaves:
        assert name in shapes[mod]


def test_prngkey():
    import deeperwin_b200 as dpe
    k = dpe.PRNGKey((5 << 32) + 7)
    assert k.dtype == torch.uint32 and k.tolist() == [5, 7]
