"""Parameter gradient and KFAC statistics (dpe_param_gradient) against the fp64 oracle (oracle/gradient.py, itself pinned to the
reference's jax.value_and_grad(total_energy) in tests/test_reference_pin.py).  Tolerance per leaf / factor: 1e-4 of its largest
entry, or 16x (oracle/parity_rule.HARD_FACTOR) the error of the SAME oracle run in fp32 on the CPU where that is larger (randomly placed walkers include
near-singular determinants whose backward pass amplifies fp32 round-off for any fp32 implementation; the factor covers the
GPU's FMA chains, measured at ~3x the CPU BLAS round-off in profiles/r02_parity_table.md, and the different summation order).  Structural mistakes show up as O(1) errors."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
from test_gpu_parity import SMALL, make  # noqa: E402


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _h_chain(n):
    import deeperwin_b200 as dpe
    return dpe.PhysicalConfig(name=f"HChain{n}", R=[[1.8 * k, 0.0, 0.0] for k in range(n)], Z=[1] * n, n_electrons=n, n_up=n // 2,
                              el_ion_mapping=list(range(0, n, 2)) + list(range(1, n, 2)))


@pytest.mark.parametrize("name,small,B", [("LiH", True, 24), ("LiH", False, 16), ("B", True, 12), ("N2", False, 6), ("HChain6", True, 8),
                                          ("N2", False, 160),           # >= 1024 (walker, electron) rows: the wide products run on the tensor cores
                                          ("Allene_TinyMol", True, 4),  # 22 electrons: inverse kernel with 32 register rows
                                          ("Benzene", True, 3),         # 42 electrons: two warps per matrix, 48 register rows
                                          ("HChain20", True, 3),        # 20 ions: more than one ion block in the envelope backward
                                          ("HChain50", True, 2)])       # 50 electrons: 64 register rows
def test_gradient_and_kfac_match_oracle(name, small, B):
    from oracle import gradient as og
    custom = _h_chain(int(name[6:])) if name.startswith("HChain") and name != "HChain6" else None
    phys, d, p32, p64, R, r, eng = make(name, B, small=small, phys=custom)
    g = torch.Generator().manual_seed(5)
    cot = (torch.randn(B, generator=g) / B).float()
    flat, lp = eng.param_gradient(r.cuda(), cot.cuda(), with_kfac=True)
    ref = og.param_gradient(p64, d, r.double(), R.double(), phys.Z, cot.double())
    ref32 = og.param_gradient(p32, d, r, R, phys.Z, cot)
    floor = max(_rel(ref32[mod][leaf], ref[mod][leaf]) for mod, leaf in eng.leaves)
    for (mod, leaf), (off, size, rows, cols) in zip(eng.leaves, eng.leaf_shapes):
        got = flat[off:off + size].reshape(ref[mod][leaf].shape)
        err = _rel(got, ref[mod][leaf])
        assert err < max(1e-4, 16 * floor), (mod, leaf, err, floor)
    fac = og.kfac_factors(p64, d, r.double(), R.double(), phys.Z)
    fac32 = og.kfac_factors(p32, d, r, R, phys.Z)
    floor_g = max(_rel(fac32[k][1], fac[k][1]) for k in fac)
    kf = flat[eng.n_params:]
    layers = eng.kfac_layers()
    assert {l[0] for l in layers} == set(fac)
    for lname, din, dout, hb, rpw, a_off, g_off in layers:
        A_ref, G_ref, rpw_ref = fac[lname]
        assert rpw == rpw_ref, lname
        A = kf[a_off:a_off + (din + hb) ** 2].reshape(din + hb, din + hb)
        G = kf[g_off:g_off + dout * dout].reshape(dout, dout)
        assert _rel(A, A_ref) < 1e-5, (lname, "A", _rel(A, A_ref))
        assert _rel(G, G_ref) < max(1e-4, 16 * floor_g), (lname, "G", _rel(G, G_ref), floor_g)
    # the same pass returns log psi^2
    assert torch.allclose(lp, eng.log_psi_sqr(r.cuda())[1], rtol=2e-6, atol=0)


def test_chunked_gradient_equals_single_pass_and_public_api():
    """A workspace smaller than the batch -> chunks with accumulation; build_value_and_grad_func mirrors loss_function.py:75-154."""
    import deeperwin_b200 as dpe
    from oracle import gradient as og, model as om
    cfg = dpe.Configuration(physical=dict(name="LiH"))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=2, device="cuda:0")
    eng = f.engine
    st = dpe.MCMCState.initialize_around_nuclei(96, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1), device="cuda:0")
    st = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=10, initialization="gaussian")).run_inter_steps(f, st, params, 2, 2, fixed)
    cot = torch.randn(96, device="cuda") / 96
    eng.set_params(params); eng.set_geometry(st.R, st.Z)
    full, _ = eng.param_gradient(st.r, cot, with_kfac=True)
    cap = eng.workspace_cap
    try:
        eng.workspace_cap = int(eng.lib.dpe_gradient_workspace_bytes(eng.handle, 40))          # chunks of 40, 40, 16 walkers
        eng._ws = None
        chunked, _ = eng.param_gradient(st.r, cot, with_kfac=True)
    finally:
        eng.workspace_cap = cap
        eng._ws = None
    scale = full.abs().max()
    assert float((full - chunked).abs().max() / scale) < 2e-6          # chunk sums are re-associated: equal to fp32 round-off
    # public API: value_and_grad
    gle = dpe.build_local_energy(f, forward_lap=True)
    vag = dpe.build_value_and_grad_func(f, gle, dpe.ClippingConfig(), with_kfac_statistics=True)
    (loss, (cs, aux)), grads = vag(params, dpe.init_clipping_state(), (2, 2), st.build_batch(fixed))
    d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3)
    p64 = {m: {k: v.double().cpu() for k, v in l.items()} for m, l in params.items()}
    args = (d, st.r.double().cpu(), st.R.double().cpu(), phys.Z, aux["E_loc_clipped"].double().cpu())
    ref = og.loss_gradient_from_energies(p64, *args)
    p32 = {m: {k: v.float().cpu() for k, v in l.items()} for m, l in params.items()}
    ref32 = og.loss_gradient_from_energies(p32, args[0], st.r.cpu(), st.R.cpu(), phys.Z, aux["E_loc_clipped"].cpu())
    floor = max(_rel(ref32[m][k], v) for m, l in ref.items() for k, v in l.items())
    for m, leaves in ref.items():
        for k, v in leaves.items():
            assert grads[m][k].shape == params[m][k].shape
            assert _rel(grads[m][k], v) < max(1e-4, 16 * floor), (m, k, _rel(grads[m][k], v), floor)
    assert abs(float(loss) - float(aux["E_mean_clipped"])) == 0.0 and "kfac" in aux and len(aux["kfac"]) == 32


@pytest.mark.parametrize("name,B,nd", [("LiH", 16, 4), ("N2", 6, 4), ("B", 10, 3)])
def test_tao_models_gradient_of_the_embedding_and_kfac(name, B, nd):
    """TAO orbital head (geometry cache held fixed): gradient with respect to the embedding parameters and the Kronecker factors of the
    embedding layers, against the fp64 oracle (autograd through oracle.model.orbitals_tao)."""
    from oracle import gradient as og, model as om
    phys, d, p32, p64, R, r, eng = make(name, B, n_dets=nd, use_taos=True)
    tao32 = om.cast_tao_cache(om.make_tao_cache(d, seed=9), torch.float32)
    tao64 = om.cast_tao_cache(tao32, torch.float64)
    eng.set_tao_cache({k: [t.cuda() for t in v] for k, v in tao32.items()})
    g = torch.Generator().manual_seed(6)
    cot = (torch.randn(B, generator=g) / B).float()
    flat, lp = eng.param_gradient(r.cuda(), cot.cuda(), with_kfac=True)
    ref = og.param_gradient(p64, d, r.double(), R.double(), phys.Z, cot.double(), tao64)
    ref32 = og.param_gradient(p32, d, r, R, phys.Z, cot, tao32)
    assert all(not m.startswith("wf/~/orbitals") for m, _ in eng.leaves)          # no per-walker orbital parameters in a TAO model
    floor = max(_rel(ref32[mod][leaf], ref[mod][leaf]) for mod, leaf in eng.leaves)
    for (mod, leaf), (off, size, rows, cols) in zip(eng.leaves, eng.leaf_shapes):
        err = _rel(flat[off:off + size].reshape(ref[mod][leaf].shape), ref[mod][leaf])
        assert err < max(1e-4, 16 * floor), (mod, leaf, err, floor)
    fac = og.kfac_factors(p64, d, r.double(), R.double(), phys.Z, tao64)
    fac32 = og.kfac_factors(p32, d, r, R, phys.Z, tao32)
    floor_g = max(_rel(fac32[k][1], fac[k][1]) for k in fac)
    kf = flat[eng.n_params:]
    layers = eng.kfac_layers()
    assert {l[0] for l in layers} == set(fac) and not any("orbitals" in l[0] for l in layers)
    for lname, din, dout, hb, rpw, a_off, g_off in layers:
        A = kf[a_off:a_off + (din + hb) ** 2].reshape(din + hb, din + hb)
        G = kf[g_off:g_off + dout * dout].reshape(dout, dout)
        assert _rel(A, fac[lname][0]) < 1e-5, (lname, "A")
        assert _rel(G, fac[lname][1]) < max(1e-4, 16 * floor_g), (lname, "G", _rel(G, fac[lname][1]), floor_g)
    assert torch.allclose(lp, eng.log_psi_sqr(r.cuda())[1], rtol=2e-6, atol=0)


_PRODUCTS_SCRIPT = """
import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_gpu_parity import make
phys, d, p32, p64, R, r, eng = make("N2", 160)
cot = (torch.randn(160, generator=torch.Generator().manual_seed(5)) / 160).float()
flat, lp = eng.param_gradient(r.cuda(), cot.cuda(), with_kfac=True)
np.save(sys.argv[1], flat.cpu().numpy())
"""


def test_tensor_core_products_of_the_gradient_pass_match_the_fp32_core_products(tmp_path):
    """The gradient + KFAC pass of N2 x 160 walkers with its wide products and dx = dz W^T on tcgen05 (default) against the same pass in a fresh
    process with DPE_ATB_TC=0 DPE_GEMM_NT_TC=0 (the switches are read once per process): identical inputs up to those products, so the
    difference is their own 3xTF32 rounding, propagated through the linear backward chain -- far below the oracle tolerance."""
    import os, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    out = {}
    for tag, env in (("tc", {}), ("fp32", {"DPE_ATB_TC": "0", "DPE_GEMM_NT_TC": "0"})):
        f = tmp_path / f"{tag}.npy"
        subprocess.run([sys.executable, "-c", _PRODUCTS_SCRIPT, str(f)], cwd=root, env={**os.environ, **env}, check=True, timeout=600)
        out[tag] = torch.from_numpy(np.load(f))
    phys, d, p32, p64, R, r, eng = make("N2", 4)
    worst = 0.0
    for (mod, leaf), (off, size, rows, cols) in zip(eng.leaves, eng.leaf_shapes):
        worst = max(worst, _rel(out["tc"][off:off + size], out["fp32"][off:off + size]))
    kf = slice(eng.n_params, None)
    for lname, din, dout, hb, rpw, a_off, g_off in eng.kfac_layers():
        for o, n in ((a_off, (din + hb) ** 2), (g_off, dout * dout)):
            worst = max(worst, _rel(out["tc"][kf][o:o + n], out["fp32"][kf][o:o + n]))
    assert not torch.equal(out["tc"], out["fp32"])          # the switch did change the path
    print("tensor-core vs FP32-core products, worst relative difference of a leaf / factor:", worst)
    assert worst < 1e-4, worst          # measured 3.6e-5 (leaves whose terms cancel: both paths carry fp32 round-off of the summands, in different orders)
