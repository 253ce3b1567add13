"""A small pass through every kernel family for compute-sanitizer (memcheck / racecheck): E_loc + forward on LiH / N2 / ethene (N > 16 path),
Metropolis steps with every proposal, gradient + KFAC pass (small batches: FP32-core products; 160 walkers: tensor-core products), TAO head, XLA shim is covered by its own test."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import deeperwin_b200 as dpe

small = dict(embedding=dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], n_hidden_el_ions=[4], emb_dim=8), orbitals=dict(n_determinants=3))
for mol, B, model_kw in (("LiH", 8, {}), ("N2", 6, {}), ("Ethene", 3, small), ("Benzene", 2, small)):
    cfg = dpe.Configuration(physical=dict(name=mol), model=model_kw)
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1, device="cuda:0")
    gle = dpe.build_local_energy(f, forward_lap=True)
    st = dpe.MCMCState.initialize_around_nuclei(B, phys, "exponential", "el_ion_mapping", dpe.PRNGKey(3), device="cuda:0")
    for prop in ("normal", "cauchy", "normal_one_el", "local", "local_one_el", "langevin"):
        mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=2, proposal=dict(name=prop)))
        for _ in range(3):                                   # third call replays the captured graph
            st = mc.run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
    e = gle(params, (phys.n_up, phys.n_dn), *st.build_batch(fixed))
    vag = dpe.build_value_and_grad_func(f, gle, dpe.ClippingConfig(), with_kfac_statistics=True)
    (loss, (cs, aux)), grads = vag(params, dpe.init_clipping_state(), (phys.n_up, phys.n_dn), st.build_batch(fixed))
    for path in (0, 1):
        f.engine.set_gemm_path(path) if path == 0 or f.engine.lib.dpe_get_gemm_path(f.engine.handle) == 0 else None
        f.engine.local_energy(st.r)
    f.engine.set_gemm_path(1)
    torch.cuda.synchronize()
    print(mol, "E_mean", float(e.mean()), "loss", float(loss), "finite grads", all(torch.isfinite(v).all() for l in grads.values() for v in l.values()))
# TAO orbital head (geometry cache = random stand-ins of the right shapes): forward, E_loc, gradient of the embedding
import math
cfg = dpe.Configuration(physical=dict(name="LiH"), model=dict(orbitals=dict(envelope_orbitals=None, transferable_atomic_orbitals=dict(name="taos"), n_determinants=4)))
phys = cfg.physical
f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1, device="cuda:0")
emb = cfg.model.embedding.n_hidden_one_el[-1]
g = torch.Generator().manual_seed(3)
tao = {"backflows": [(torch.randn(len(phys.Z), n, 2, 4, emb, generator=g) / math.sqrt(emb)).cuda() for n in (phys.n_up, phys.n_dn)],
       "exponents": [(0.5 + torch.rand(len(phys.Z), n, 2, 4, generator=g)).cuda() for n in (phys.n_up, phys.n_dn)]}
fixed = dict(fixed or {}, cache=dict(taos=tao))
st = dpe.MCMCState.initialize_around_nuclei(8, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(3), device="cuda:0")
mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=2))
for _ in range(3):
    st = mc.run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
gle = dpe.build_local_energy(f, forward_lap=True)
vag = dpe.build_value_and_grad_func(f, gle, dpe.ClippingConfig(), with_kfac_statistics=True)
(loss, _), grads = vag(params, dpe.init_clipping_state(), (phys.n_up, phys.n_dn), st.build_batch(fixed))
torch.cuda.synchronize()
print("LiH TAO loss", float(loss), "finite grads", all(torch.isfinite(v).all() for l in grads.values() for v in l.values()))
# >= 1024 (walker, electron) rows: the wide gradient / KFAC products and dx = dz W^T on the tensor cores (k_atb_prep, launch_atb_tc, dense_gemm_t),
# the thread-per-pair backward of the pair stream (k_bw_pair_rows)
cfg = dpe.Configuration(physical=dict(name="N2"))
phys = cfg.physical
f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1, device="cuda:0")
st = dpe.MCMCState.initialize_around_nuclei(160, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(3), device="cuda:0")
f.engine.set_params(params); f.engine.set_geometry(st.R, st.Z)
flat, lp = f.engine.param_gradient(st.r, torch.full((160,), 1.0 / 160, device="cuda"), with_kfac=True)
torch.cuda.synchronize()
print("N2 x 160 gradient + KFAC finite", bool(torch.isfinite(flat).all()))
print("done")
