"""The oracle has no reference golden vectors to lean on (PARITY UNPINNED, see oracle/model.py), so it is pinned
by independent definitions and invariants: three Laplacians, antisymmetry, closed-form potentials and an
analytic wavefunction."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import model as om

GOLD = Path(__file__).parent / "golden"
SMALL = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)


def lih(small=True, **kw):
    R = torch.tensor([[0, 0, 0], [3.015, 0, 0]], dtype=torch.float64)
    Z = [3, 1]
    d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3, **(SMALL if small else {}), **kw)
    return R, Z, d, [0, 1, 0, 0]


def walkers(R, mapping, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return R[torch.tensor(mapping)][None] + torch.randn(B, len(mapping), 3, generator=g, dtype=torch.float64)


def test_three_laplacians_agree():
    R, Z, d, m = lih()
    params = om.init_params(d, seed=1, bias_scale=0.1, envelope_jitter=0.5)
    r = walkers(R, m, 3)
    ek_h, g_h, lap_h = om.kinetic_energy_hessian(params, d, r, R, Z)
    fl = om.forward_laplacian(params, d, r, R, Z)
    ek_j = om.kinetic_energy_jvp_loop(params, d, r[:2], R, Z)           # hamiltonian.py:234-267
    assert torch.allclose(fl["grad"], g_h, rtol=1e-9, atol=1e-10)
    assert torch.allclose(fl["lap"], lap_h, rtol=1e-9, atol=1e-9)
    assert torch.allclose(fl["E_kin"], ek_h, rtol=1e-9, atol=1e-9)
    assert torch.allclose(ek_j, ek_h[:2], rtol=1e-9, atol=1e-9)
    assert torch.allclose(fl["logpsi2"], om.log_psi_sqr(params, d, r, R, Z)[1], rtol=1e-12)


def test_default_width_model_forward_laplacian_matches_hessian():
    R, Z, d, m = lih(small=False)
    params = om.init_params(d, seed=2, bias_scale=0.05, envelope_jitter=0.3)
    r = walkers(R, m, 2, seed=3)
    ek_h, g_h, _ = om.kinetic_energy_hessian(params, d, r, R, Z)
    fl = om.forward_laplacian(params, d, r, R, Z)
    assert torch.allclose(fl["E_kin"], ek_h, rtol=1e-8, atol=1e-8)
    assert torch.allclose(fl["grad"], g_h, rtol=1e-8, atol=1e-9)


def test_antisymmetry():
    """Swapping two same-spin electrons leaves log psi^2 and E_loc unchanged and flips the phase."""
    R, Z, d, m = lih()
    params = om.init_params(d, seed=4, bias_scale=0.1, envelope_jitter=0.5)
    r = walkers(R, m, 4, seed=1)
    r_sw = r.clone()
    r_sw[:, [0, 1]] = r[:, [1, 0]]
    a, b = om.forward_laplacian(params, d, r, R, Z), om.forward_laplacian(params, d, r_sw, R, Z)
    assert torch.allclose(a["logpsi2"], b["logpsi2"], rtol=1e-10)
    assert torch.allclose(a["E_loc"], b["E_loc"], rtol=1e-8, atol=1e-8)
    assert torch.all((a["phase"] - b["phase"]).abs() == math.pi)


def test_potential_energy_closed_forms():
    R = torch.tensor([[0.0, 0, 0], [2.0, 0, 0]], dtype=torch.float64)
    r = torch.tensor([[[0.0, 1.0, 0], [2.0, 0, 2.0]]], dtype=torch.float64)
    d12 = math.sqrt(4 + 1 + 4)
    expect = 1 / d12 - (3 / 1.0 + 1 / math.sqrt(5) + 3 / math.sqrt(8) + 1 / 2.0) + 3 * 1 / 2.0
    assert abs(om.potential_energy(r, R, [3, 1]).item() - expect) < 1e-12
    # single ion: no ion-ion term
    assert abs(om.potential_energy(r, R[:1], [2]).item() - (1 / d12 - 2 / 1.0 - 2 / math.sqrt(8))) < 1e-12


def test_analytic_helium_like_wavefunction():
    """one nucleus, U=D=1, zero embedding weights, constant backflow, softplus(alpha)=Z
    => psi ~ exp(-Z (r1 + r2)) and E_loc = -Z^2 + 1/r12 for every walker."""
    Zc = 2
    R = torch.zeros(1, 3, dtype=torch.float64)
    d = om.ModelDims(n_el=2, n_up=1, n_ion=1, Z_max=2, n_iterations=1, n_hidden_one_el=[8], n_hidden_two_el=[], emb_dim=4, n_dets=1)
    params = om.init_params(d, seed=0)
    for mod, leaves in params.items():
        for k in leaves:
            if k in ("w", "b", "embeddings"):
                leaves[k].zero_()
    params[f"{om.EMB}/h_el_0/linear_0"]["b"].fill_(0.7)
    # full_det 2x2 with rows (up, dn) and columns (orb 0, orb 1): make it diagonal so that det = phi(r1) phi(r2)
    w_up = torch.zeros(8, 2, dtype=torch.float64); w_up[:, 0] = 0.3
    w_dn = torch.zeros(8, 2, dtype=torch.float64); w_dn[:, 1] = 0.3
    params[f"{om.ORB}/bf_up/linear_0"]["w"] = w_up
    params[f"{om.ORB}/bf_dn/linear_0"]["w"] = w_dn
    alpha = math.log(math.expm1(Zc))           # softplus(alpha) = Z
    for k in ("alpha_up", "alpha_dn"):
        params[om.ORB][k].fill_(alpha)
    r = torch.randn(5, 2, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    out = om.forward_laplacian(params, d, r, R, [Zc])
    r12 = (r[:, 0] - r[:, 1]).norm(dim=-1)
    assert torch.allclose(out["E_loc"], -Zc ** 2 + 1 / r12, rtol=1e-7, atol=1e-7)
    assert torch.allclose(out["logpsi2"], 2 * (-Zc * r.norm(dim=-1).sum(-1)) + out["logpsi2"][0] + 2 * Zc * r[0].norm(dim=-1).sum(), rtol=1e-9)


def test_tao_orbitals_three_definitions_and_antisymmetry():
    """SURVEY.md 8 a18 (transferable_atomic_orbitals.py:287-349, cache branch): the explicit forward-Laplacian through the TAO
    orbitals equals the autodiff Hessian of the literal einsum restatement, for even and odd spin blocks."""
    for n_el, n_up, n_ion, zs in ((4, 2, 2, [3, 1]), (5, 3, 1, [5])):
        d = om.ModelDims(n_el=n_el, n_up=n_up, n_ion=n_ion, Z_max=max(zs), use_taos=True, **{**SMALL, "n_dets": 2})
        params = om.init_params(d, seed=3, bias_scale=0.1)
        assert not any("orbitals" in k for k in params)           # no per-walker orbital parameters
        tao = om.make_tao_cache(d, seed=5)
        assert tao["backflows"][0].shape == (n_ion, n_up, 2, 2, 16) and tao["exponents"][1].shape == (n_ion, n_el - n_up, 2, 2)
        R = torch.zeros(n_ion, 3, dtype=torch.float64)
        R[-1, 0] = 3.015 if n_ion > 1 else 0.0
        g = torch.Generator().manual_seed(n_el)
        r = torch.randn(3, n_el, 3, generator=g, dtype=torch.float64)
        fl = om.forward_laplacian(params, d, r, R, zs, tao=tao)
        ek, grad, lap = om.kinetic_energy_hessian(params, d, r, R, zs, tao)
        assert torch.allclose(fl["logpsi2"], om.log_psi_sqr(params, d, r, R, zs, tao)[1], rtol=1e-12)
        assert torch.allclose(fl["grad"], grad, rtol=1e-9, atol=1e-10)
        assert torch.allclose(fl["lap"], lap, rtol=1e-9, atol=1e-8)
        assert torch.allclose(fl["E_kin"], ek, rtol=1e-9, atol=1e-9)
        r_sw = r.clone()
        r_sw[:, [0, 1]] = r[:, [1, 0]]                            # two spin-up electrons
        sw = om.forward_laplacian(params, d, r_sw, R, zs, tao=tao)
        assert torch.allclose(sw["logpsi2"], fl["logpsi2"], rtol=1e-10) and torch.allclose(sw["E_loc"], fl["E_loc"], rtol=1e-8, atol=1e-8)
        assert torch.all((sw["phase"] - fl["phase"]).abs() > 3.0)
        # the different-spin exponent slice is really used: changing it moves the result (guards the same/diff bookkeeping)
        tao2 = {"backflows": tao["backflows"], "exponents": [t.clone() for t in tao["exponents"]]}
        tao2["exponents"][0][:, :, 1] *= 1.1
        assert not torch.allclose(om.log_psi_sqr(params, d, r, R, zs, tao2)[1], fl["logpsi2"], rtol=1e-6)
        # ... while the different-spin BACKFLOW slice is not (transferable_atomic_orbitals.py:255-260 uses b_same twice)
        tao3 = {"backflows": [t.clone() for t in tao["backflows"]], "exponents": tao["exponents"]}
        tao3["backflows"][0][:, :, 1] += 1.0
        assert torch.allclose(om.log_psi_sqr(params, d, r, R, zs, tao3)[1], fl["logpsi2"], rtol=1e-12)


@pytest.mark.parametrize("name", ["LiH_small", "LiH", "LiH_tao"])
def test_golden_fixture_regression(name):
    g = np.load(GOLD / f"model_{name}.npz")
    kw = SMALL if name.endswith("small") else (dict(n_dets=4, use_taos=True) if name.endswith("tao") else {})
    d = om.ModelDims(n_el=g["r"].shape[1], n_up=int(g["n_up"]), n_ion=len(g["Z"]), Z_max=int(g["Z"].max()), **kw)
    params = om.cast_params(om.cast_params(om.init_params(d, seed=int(g["seed"]), bias_scale=float(g["bias_scale"]),
                                                          envelope_jitter=float(g["envelope_jitter"])), torch.float32), torch.float64)
    tao = om.cast_tao_cache(om.cast_tao_cache(om.make_tao_cache(d, seed=int(g["seed"])), torch.float32), torch.float64) if d.use_taos else None
    out = om.forward_laplacian(params, d, torch.from_numpy(g["r"]).double(), torch.from_numpy(g["R"]).double(), g["Z"].tolist(), tao=tao)
    for k in ("logpsi2", "grad", "E_kin", "E_pot", "E_loc"):
        assert np.allclose(out[k].numpy(), g[k], rtol=1e-9, atol=1e-9), k
    assert np.array_equal(out["phase"].numpy(), g["phase"])


def test_param_tree_names_and_shapes():
    d = om.ModelDims(n_el=14, n_up=7, n_ion=2, Z_max=7)
    sh = om.param_shapes(d)
    assert sh["wf/~/input/h_ion"]["embeddings"] == (7, 32)
    assert sh[f"{om.EMB}/h_el_0/linear_0"]["w"] == (3 * 8 + 32 + 4, 256)
    assert sh[f"{om.EMB}/h_el_1/linear_0"]["w"] == (832, 256)
    assert sh[f"{om.EMB}/symm_features_0/convolutional_features/h_ion_map/linear_0"]["w"] == (32, 4)
    assert sh[f"{om.ORB}/bf_up/linear_0"]["w"] == (256, 32 * 14)
    assert f"{om.EMB}/h_same_3/linear_0" not in sh
    n = sum(int(np.prod(s)) for leaves in sh.values() for s in leaves.values())
    assert 0.9e6 < n < 0.96e6        # 0.93 M parameters for N2


@pytest.mark.parametrize("name", ["LiH", "N2", "LiH_small", "B_small", "Ethene_small"])
def test_oracle_reproduces_reference_fixtures(name):
    """tests/golden/reference_*.npz hold outputs of the REFERENCE'S OWN CODE (unmodified modules executed under tests/ref_shim,
    tests/golden/make_reference_golden.py).  The oracle reproduces them to round-off: this is what pins oracle/model.py on machines
    without /root/reference (tests/test_reference_pin.py does the comparison live where the tree exists)."""
    import numpy as np
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / f"reference_{name}.npz")
    kw = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3) if bool(g["small"]) else {}
    d = om.ModelDims(n_el=g["r"].shape[1], n_up=int(g["n_up"]), n_ion=len(g["Z"]), Z_max=int(g["Z"].max()), **kw)
    p32 = om.cast_params(om.init_params(d, seed=int(g["seed"]), bias_scale=float(g["bias_scale"]), envelope_jitter=float(g["envelope_jitter"])), torch.float32)
    chk = float(sum(v.double().abs().sum() for l in p32.values() for v in l.values()))
    assert abs(chk - float(g["param_checksum"])) <= 1e-9 * chk
    out = om.forward_laplacian(om.cast_params(p32, torch.float64), d, torch.from_numpy(g["r"]).double(), torch.from_numpy(g["R"]).double(), g["Z"].tolist())
    assert np.allclose(out["logpsi2"].numpy(), g["logpsi2"], rtol=1e-12, atol=1e-12)
    assert np.allclose(out["E_loc"].numpy(), g["E_loc"], rtol=1e-9, atol=1e-9)
    assert np.allclose(out["E_pot"].numpy(), g["E_pot"], rtol=1e-12)
    assert np.array_equal(out["phase"].numpy() > 1, g["phase"] > 1)
