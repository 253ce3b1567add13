"""GPU debugging aid: runs the CUDA path on a few walkers and prints, stage by stage, the max error of the
intermediates in the workspace against the fp64 oracle (test infrastructure: imports oracle/).
Usage (on a GPU box): python tools/stage_report.py [LiH_small|LiH|N2|Benzene] [B] [gemm_path]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np
import torch

from oracle import model as om
from deeperwin_b200.engine import Engine
from deeperwin_b200._lib import MODE_FORWARD, MODE_LAPLACIAN
from deeperwin_b200.configuration import PhysicalConfig


def make_case(name):
    small = name.endswith("_small")
    phys = PhysicalConfig(name=name.replace("_small", ""))
    kw = dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3) if small else {}
    d = om.ModelDims(n_el=phys.n_electrons, n_up=phys.n_up, n_ion=len(phys.Z), Z_max=max(phys.Z), **kw)
    return phys, d


def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "LiH_small"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    phys, d = make_case(name)
    params64 = om.init_params(d, seed=3, bias_scale=0.1, envelope_jitter=0.5)
    params32 = om.cast_params(params64, torch.float32)
    params64 = om.cast_params(params32, torch.float64)
    g = torch.Generator().manual_seed(7)
    R = torch.tensor(phys.R, dtype=torch.float32)
    r32 = (R[torch.tensor(phys.el_ion_mapping)][None] + torch.randn(B, d.n_el, 3, generator=g)).float()
    r64 = r32.double()
    ref = om.forward_laplacian(params64, d, r64, R.double(), phys.Z, return_intermediates=True)
    inter = ref["inter"]

    eng = Engine(n_el=d.n_el, n_up=d.n_up, n_ion=d.n_ion, n_iterations=d.n_iterations, n_hidden_one_el=d.n_hidden_one_el,
                 n_hidden_two_el=d.n_hidden_two_el, emb_dim=d.emb_dim, n_ion_features=d.n_ion_features, n_dets=d.n_dets,
                 z_min=d.Z_min, z_max=d.Z_max)
    eng.set_gemm_path(path)
    eng.set_params({m: {k: v.cuda() for k, v in l.items()} for m, l in params32.items()})
    eng.set_geometry(R, phys.Z)
    e_loc, aux = eng.local_energy(r32.cuda(), with_aux=True)
    torch.cuda.synchronize()
    N, C, K = d.n_el, 3 * d.n_el + 2, 3 * d.n_el
    nit = d.n_iterations
    ldx = eng.ldx(B, MODE_LAPLACIAN)
    print(f"case {name} B={B} N={N} C={C} ldx={ldx} gemm_path={path}")
    for it in range(nit):
        pw = eng.ws_view(f"pw{it}", B, MODE_LAPLACIAN, (B, N, N, 3, d.emb_dim))
        print(f"  pair w_{it}: {rel(pw, inter[f'w_{it}']):.2e}")
    for it in range(nit):
        dE = d.d_eion_in(it)
        ei = eng.ws_view(f"ei{it}", B, MODE_LAPLACIAN, (B, N, 5, dE))
        f = inter[f"f_{it}"]
        ce = f[..., f.shape[-1] - dE:]
        refei = torch.stack([ce[:, :, 0]] + [torch.stack([ce[:, i, 1 + 3 * i + a] for i in range(N)], 1) for a in range(3)] + [ce[:, :, C - 1]], 2)
        print(f"  el-ion conv_{it}: {rel(ei, refei):.2e}")
    last = nit - 1
    xin = eng.ws_view("x1" if nit % 2 == 0 else "x0", B, MODE_LAPLACIAN, (B, N, C, ldx))
    xout = eng.ws_view("x0" if nit % 2 == 0 else "x1", B, MODE_LAPLACIAN, (B, N, C, ldx))
    f = inter[f"f_{last}"]
    din = d.d_one_in(last)
    dE = d.d_eion_in(last)
    print(f"  h_in (it {last}): {rel(xin[..., :din], f[..., :din]):.2e}")
    print(f"  conv_ee (it {last}): {rel(xin[..., din:din + d.emb_dim], f[..., 3 * din:3 * din + d.emb_dim]):.2e}")
    print(f"  conv_eI (it {last}): {rel(xin[..., din + d.emb_dim:din + d.emb_dim + dE], f[..., 3 * din + d.emb_dim:]):.2e}")
    hm = eng.ws_view("hm", B, MODE_LAPLACIAN, (B, N, C, d.emb_dim))
    print(f"  hm (it {last}): {rel(hm, inter[f'hm_{last}']):.2e}")
    dout = d.n_hidden_one_el[last]
    print(f"  h_out (it {last}): value {rel(xout[:, :, 0, :dout], inter[f'h_{last}'][:, :, 0]):.2e} "
          f"tangents {rel(xout[:, :, 1:C - 1, :dout], inter[f'h_{last}'][:, :, 1:C - 1]):.2e} "
          f"lap {rel(xout[:, :, C - 1, :dout], inter[f'h_{last}'][:, :, C - 1]):.2e}")
    mo = eng.ws_view("mo", B, MODE_LAPLACIAN, (B, N, C, d.n_dets * N))
    print(f"  mo: value {rel(mo[:, :, 0], inter['mo'][:, :, 0]):.2e} tangents {rel(mo[:, :, 1:C - 1], inter['mo'][:, :, 1:C - 1]):.2e} "
          f"lap {rel(mo[:, :, C - 1], inter['mo'][:, :, C - 1]):.2e}")
    det = eng.ws_view("det", B, MODE_LAPLACIAN, (B, d.n_dets, K + 3))
    print(f"  det: logdet {rel(det[..., 0], ref['logdet_d']):.2e} sign_eq {bool((det[..., 1].cpu().double() == ref['sign_d']).all())} "
          f"lap'_d {rel(det[..., 2], inter['lap_d'] + (inter['g_d'] ** 2).sum(-1)):.2e} g_d {rel(det[..., 3:], inter['g_d']):.2e}")
    print(f"  logpsi2 {rel(aux['log_psi_sqr'], ref['logpsi2']):.2e}  grad {rel(aux['grad'], ref['grad']):.2e}  "
          f"E_kin {rel(aux['E_kin'], ref['E_kin']):.2e}  E_pot {rel(aux['E_pot'], ref['E_pot']):.2e}")
    print("  per-walker E_kin rel err", ((aux["E_kin"].double().cpu() - ref["E_kin"]).abs() / ref["E_kin"].abs()).tolist())
    dd = det.double().cpu()
    lp_ref = inter["lap_d"] + (inter["g_d"] ** 2).sum(-1)
    print("  lap'_d per-element rel err max", ((dd[..., 2] - lp_ref).abs() / lp_ref.abs()).max().item(),
          " g_d per-det rel err max", ((dd[..., 3:] - inter["g_d"]).abs().amax(-1) / inter["g_d"].abs().amax(-1)).max().item())
    err = ((e_loc.double().cpu() - ref["E_loc"]).abs() / ref["E_loc"].abs()).max().item()
    print(f"  E_loc max rel err {err:.2e}   (E_loc ref {ref['E_loc'][:3].tolist()})")
    ph, lp = eng.log_psi_sqr(r32.cuda())
    print(f"  forward-only: logpsi2 {rel(lp, ref['logpsi2']):.2e}  |dlogpsi2| max {(lp.double().cpu() - ref['logpsi2']).abs().max().item():.2e} "
          f"phase_eq {bool(((ph.cpu() > 1) == (ref['phase'] > 1)).all())}")
    print(f"  launches so far: {eng.launch_count()}")
    mcmc_report(eng, phys, d, params64, r32, R)


def mcmc_report(eng, phys, d, params64, r32, R):
    import ctypes as C
    from oracle import threefry, mcmc as omc
    from deeperwin_b200 import mcmc as gm
    from deeperwin_b200.configuration import MCMCConfigOptimization
    B, N = r32.shape[0], d.n_el
    keys = threefry.split(threefry.prng_key(1234), B)
    nk, noise, thr = threefry.mcmc_step_randoms(keys, N)
    kd = torch.from_numpy(keys.view(np.int32)).cuda()
    nk_d = torch.empty_like(kd); noise_d = torch.empty(B, N, 3, device="cuda"); thr_d = torch.empty(B, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = eng.lib.dpe_threefry_mcmc_randoms(p(kd), B, N, p(nk_d), p(noise_d), p(thr_d), None)
    torch.cuda.synchronize()
    print(f"  rng: rc={rc} new_keys_eq {bool((nk_d.cpu().numpy().view(np.uint32) == nk).all())} thr_eq {bool((thr_d.cpu().numpy() == thr).all())} "
          f"noise max abs diff {np.abs(noise_d.cpu().numpy() - noise).max():.2e} bit_eq {(noise_d.cpu().numpy() == noise).mean():.3f}")
    bits = gm.random_bits(gm.PRNGKey(0), 7, "cuda").cpu().numpy()
    print(f"  bits(PRNGKey(0),7) eq {bool((bits == threefry.random_bits(threefry.prng_key(0), 7)).all())}; normal(PRNGKey(0)) = {gm.normal(gm.PRNGKey(0), (1,), 'cuda').item():.8f} (jax: -0.20584226)")
    # one chain of 5 steps vs the oracle driven by the fp64 model
    def func(r):
        return om.log_psi_sqr(params64, d, torch.from_numpy(r).double(), R.double(), phys.Z)[1].float().numpy()
    st0 = omc.OracleMCMCState(r=r32.numpy().copy(), R=R.numpy(), Z=np.array(phys.Z), log_psi_sqr=-np.ones(B, np.float32) * 1000,
                              walker_age=np.zeros(B, np.int32), rng_state=keys.copy())
    cfg = MCMCConfigOptimization(n_inter_steps=5, stepsize_update_interval=2)
    ref = omc.run_mcmc_steps(func, st0, 5, max_age=cfg.max_age, stepsize_update_interval=2)
    gst = gm.MCMCState(r=r32.cuda(), R=R.cuda(), Z=torch.tensor(phys.Z, dtype=torch.int32).cuda(),
                       log_psi_sqr=-torch.ones(B, device="cuda") * 1000, walker_age=torch.zeros(B, dtype=torch.int32, device="cuda"),
                       rng_state=kd.view(torch.uint32).clone())
    class F:  # minimal callable carrying the engine, as build_log_psi_squared returns
        engine = eng
    mc = gm.MetropolisHastingsMonteCarlo(cfg)
    eng._param_sig = None
    pd = {m: {k: v.float().cuda() for k, v in l.items()} for m, l in params64.items()}
    out = mc.run_inter_steps(F, gst, pd, d.n_up, d.n_dn, {})
    torch.cuda.synchronize()
    print(f"  mcmc 5 steps: keys_eq {bool((out.rng_state.cpu().numpy().view(np.uint32) == ref.rng_state).all())} age_eq {bool((out.walker_age.cpu().numpy() == ref.walker_age).all())} "
          f"r max diff {np.abs(out.r.cpu().numpy() - ref.r).max():.2e} lp max diff {np.abs(out.log_psi_sqr.cpu().numpy() - ref.log_psi_sqr).max():.2e} "
          f"stepsize {out.stepsize.item():.6f}/{ref.stepsize:.6f} acc_rate {out.acc_rate.item():.6f}/{ref.acc_rate:.6f} step_nr {out.step_nr.item()}/{ref.step_nr} counts {mc.last_accept_counts.tolist()}")


if __name__ == "__main__":
    main()
