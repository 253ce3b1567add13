"""Where a weight-sharing step (H10 x 512 walkers, bench.py weight_sharing_block) spends its time: Metropolis inter-steps, loss + gradient + KFAC, update."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import deeperwin_b200 as dpe
dev = "cuda:0"; n_at = 10; n_geom = 16; walkers = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mapping = list(range(0, n_at, 2)) + list(range(1, n_at, 2))
phys = [dpe.PhysicalConfig(name=f"HChain{n_at}_{a:.3f}", R=[[float(a) * k, 0.0, 0.0] for k in range(n_at)], Z=[1] * n_at, n_electrons=n_at,
                           n_up=n_at // 2, el_ion_mapping=mapping) for a in np.linspace(1.2, 3.6, n_geom)]
cfg = dpe.Configuration(physical=phys[0].model_dump())
f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys[0], None, None, rng_seed=11, device=dev)
gle = dpe.build_local_energy(f, forward_lap=True)
vag = dpe.build_value_and_grad_func(f, gle, dpe.ClippingConfig(), with_kfac_statistics=True)
mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=20, initialization="gaussian"))
spin = (5, 5)
geoms = [dpe.GeometryDataStore(idx=g, physical_config=p, spin_state=spin, fixed_params=fixed, clipping_state=dpe.init_clipping_state(),
                               mcmc_state=dpe.MCMCState.initialize_around_nuclei(walkers, p, "gaussian", "el_ion_mapping", dpe.PRNGKey(g), device=dev))
         for g, p in enumerate(phys)]
flat_p = [t for l in params.values() for t in l.values()]
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {"mcmc": 0.0, "vag": 0.0, "upd": 0.0}; host = dict(acc)
def step(epoch, timed):
    g = geoms[epoch % n_geom]
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter(); e[0].record()
    g.mcmc_state = mc.run_inter_steps(f, g.mcmc_state, params, 5, 5, g.fixed_params)
    t1 = time.perf_counter(); e[1].record()
    (loss, (g.clipping_state, aux)), grads = vag(params, g.clipping_state, g.spin_state, g.mcmc_state.build_batch(g.fixed_params))
    t2 = time.perf_counter(); e[2].record()
    torch._foreach_add_(flat_p, [grads[m][k] for m, l in params.items() for k in l], alpha=-1e-4)
    t3 = time.perf_counter(); e[3].record()
    if timed:
        torch.cuda.synchronize()
        for k, a, b in (("mcmc", 0, 1), ("vag", 1, 2), ("upd", 2, 3)): acc[k] += e[a].elapsed_time(e[b])
        host["mcmc"] += (t1 - t0) * 1e3; host["vag"] += (t2 - t1) * 1e3; host["upd"] += (t3 - t2) * 1e3
for ep in range(48): step(ep, False)
torch.cuda.synchronize()
l0 = f.engine.launch_count()
for ep in range(48, 64): step(ep, True)
print("walkers", walkers, "device ms per step:", {k: round(v / 16, 3) for k, v in acc.items()}, "host ms:", {k: round(v / 16, 3) for k, v in host.items()},
      "launches/step", (f.engine.launch_count() - l0) / 16, "graph mode", f.engine.lib.dpe_get_mcmc_graph(f.engine.handle))
# untimed-sync variant: whole round without synchronisation
torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
for ep in range(64, 80): step(ep, False)
b.record(); torch.cuda.synchronize(); print("round of 16, no syncs:", a.elapsed_time(b) / 16, "ms/step")
