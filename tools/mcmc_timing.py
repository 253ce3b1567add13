"""Times run_inter_steps at the reference cadence (20 Metropolis steps per call) on N2 x 4096 walkers."""
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
st = dpe.MCMCState.initialize_around_nuclei(B, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1234), device="cuda:0")
mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=20, initialization="gaussian"))
for _ in range(3):
    st = mc.run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(5):
    st = mc.run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
e1.record(); t_enq = time.perf_counter() - t0
torch.cuda.synchronize(); t_wall = time.perf_counter() - t0
n_fwd = 5 * 21
print(f"{mol} B={B}: device {e0.elapsed_time(e1) / n_fwd:.3f} ms per forward pass ({B * n_fwd / e0.elapsed_time(e1) * 1e3:.0f} walker-steps/s), "
      f"host enqueue {1e3 * t_enq / n_fwd:.3f} ms per pass, wall {1e3 * t_wall / n_fwd:.3f} ms; acc_rate {st.acc_rate.item():.3f} stepsize {st.stepsize.item():.4f}")
