import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from oracle import model as om, parity_rule
from deeperwin_b200.engine import Engine
g = np.load("tests/golden/model_LiH_tao.npz")
d = om.ModelDims(n_el=4, n_up=int(g["n_up"]), n_ion=2, Z_max=3, n_dets=4, use_taos=True)
p32 = om.cast_params(om.init_params(d, seed=int(g["seed"]), bias_scale=float(g["bias_scale"]), envelope_jitter=float(g["envelope_jitter"])), torch.float32)
p64 = om.cast_params(p32, torch.float64)
eng = Engine(n_el=4, n_up=d.n_up, n_ion=2, n_iterations=4, n_hidden_one_el=d.n_hidden_one_el, n_hidden_two_el=d.n_hidden_two_el, emb_dim=32, n_ion_features=32, n_dets=4, z_min=1, z_max=3, use_taos=True)
eng.set_params({m: {k: v.cuda() for k, v in l.items()} for m, l in p32.items()})
eng.set_geometry(g["R"], g["Z"])
tao32 = om.cast_tao_cache(om.make_tao_cache(d, seed=int(g["seed"])), torch.float32)
tao64 = om.cast_tao_cache(tao32, torch.float64)
eng.set_tao_cache({k: [t.cuda() for t in v] for k, v in tao32.items()})
r, R, Z = torch.from_numpy(g["r"]), torch.from_numpy(g["R"]), g["Z"].tolist()
ref = om.forward_laplacian(p64, d, r.double(), R.double(), Z, tao=tao64, return_intermediates=True)
env = parity_rule.fp32_envelope(om, p32, d, r, R, Z, ref, tao32=tao32)
print("cond", ref["cond"].numpy().round(0)); print("floor lp", env["logpsi2"]); print("logdet w6", ref["logdet_d"][6], ref["sign_d"][6], "logpsi2", ref["logpsi2"][6])
for gp in (1, 0):
    eng.set_gemm_path(gp); eng.set_det_path(simt=(gp == 0))
    e, aux = eng.local_energy(r.cuda(), with_aux=True)
    lp = eng.log_psi_sqr(r.cuda())[1]
    err = parity_rule.errors(dict(logpsi2=lp, E_loc=e), ref)
    print("gemm", gp, "lp err", err["logpsi2"], "\n   E err", err["E_loc"], "\n  aux lp err", parity_rule.errors(dict(logpsi2=aux["log_psi_sqr"], E_loc=e), ref)["logpsi2"])
    B = 8; C = 14
    mo = eng.ws_view("mo", B, 1, (B, 4, C, 16))[:, :, 0].double().cpu()
    mo_ref = ref["inter"]["mo"][:, :, 0]
    print("   mo err rel to max", ((mo - mo_ref).abs().amax((1, 2)) / mo_ref.abs().amax((1, 2))).numpy())
    x = eng.ws_view("x0", B, 1, (B, 4, C, eng.ldx(B, 1)))[:, :, 0, :256].double().cpu()
    h = ref["inter"]["h_3"][:, :, 0]
    print("   h_3 err", ((x - h).abs().amax((1, 2))).numpy(), " (either x0 or x1 holds it)")
    x = eng.ws_view("x1", B, 1, (B, 4, C, eng.ldx(B, 1)))[:, :, 0, :256].double().cpu()
    print("   h_3 err", ((x - h).abs().amax((1, 2))).numpy())
