"""Times the Metropolis steps (N2 x 4096 walkers by default): public run_inter_steps at n_inter_steps = 1 and 20, CUDA-graph replay on / off."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import deeperwin_b200 as dpe

mol = sys.argv[1] if len(sys.argv) > 1 else "N2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = dpe.Configuration(physical=dict(name=mol))
phys = cfg.physical
f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1234, device="cuda:0")
st0 = dpe.MCMCState.initialize_around_nuclei(B, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1234), device="cuda:0")
for n_inter in (1, 20):
    for graph in (False, True, False, True):
        f.engine.set_mcmc_graph(graph)
        mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=n_inter, initialization="gaussian"))
        st = st0
        for _ in range(4):
            st = mc.run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
        torch.cuda.synchronize()
        reps = 40 if n_inter == 1 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(reps):
            st = mc.run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
        e1.record(); t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        n_fwd = reps * (n_inter + 1)
        print(f"{mol} B={B} n_inter={n_inter} graph={graph}: device {e0.elapsed_time(e1) / reps:.3f} ms per call = {e0.elapsed_time(e1) / n_fwd:.3f} ms per forward pass, "
              f"host enqueue {1e3 * t_enq / reps:.3f} ms per call; graph mode after: {f.engine.lib.dpe_get_mcmc_graph(f.engine.handle)}")
