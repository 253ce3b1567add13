"""One forward-Laplacian E_loc pass (N2, 4096 walkers by default) for ncu captures."""
import sys
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
gle = dpe.build_local_energy(f, forward_lap=True)
e = gle(params, (phys.n_up, phys.n_dn), st.r, st.R, st.Z, fixed)
torch.cuda.synchronize()
print("E_mean", e.mean().item())
