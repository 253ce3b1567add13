import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import deeperwin_b200 as dpe
from oracle import model as om, parity_rule
dev = torch.device("cuda:0")
cfg = dpe.Configuration(physical=dict(name="LiH")); phys = cfg.physical
f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1234, device=dev)
gle = dpe.build_local_energy(f, forward_lap=True)
state = dpe.MCMCState.initialize_around_nuclei(64, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1234), device=dev)
mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=2, initialization="gaussian"))
state = mc.run_inter_steps(f, state, params, 2, 2, fixed)
d = om.ModelDims(n_el=4, n_up=2, n_ion=2, Z_max=3)
p32 = {m: {k: v.cpu() for k, v in l.items()} for m, l in params.items()}; p64 = om.cast_params(p32, torch.float64)
r, R = state.r.cpu(), state.R.cpu()
ref = om.forward_laplacian(p64, d, r.double(), R.double(), phys.Z)
env = parity_rule.fp32_envelope(om, p32, d, r, R, phys.Z, ref)
env32 = parity_rule.fp32_envelope(om, p32, d, r, R, phys.Z, ref, n_perm=32, seed=5)
eng = f.engine
W = [46, 57, 23]
print("cond", ref["cond"][W].numpy(), "floor8", env["E_loc"][W], "floor32", env32["E_loc"][W], "E_kin", ref["E_kin"][W].numpy(), "E_pot", ref["E_pot"][W].numpy())
for gp, simt in ((1, False), (0, False), (1, True), (0, True)):
    eng.set_gemm_path(gp); eng.set_det_path(simt=simt)
    e = eng.local_energy(state.r)
    err = parity_rule.errors(dict(logpsi2=ref["logpsi2"], E_loc=e), ref)["E_loc"]
    print(f"gemm {gp} det_simt {simt}: E err {err[W]}  median {np.median(err):.2e} max/soft8 {(err/np.maximum(1e-4,2*env['E_loc'])).max():.2f} max/soft32 {(err/np.maximum(1e-4,2*env32['E_loc'])).max():.2f}")
# minimal el-ion / el-el distances of these walkers
dn = (r[:, :, None] - R[None, None]).norm(dim=-1).amin((1, 2)); 
print("min el-ion dist", dn[W].numpy(), "median", dn.median().item())
