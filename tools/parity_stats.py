"""Prints per-walker error statistics of the CUDA path vs the fp64 oracle next to the fp32 CPU oracle's own error."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
from oracle import model as om
from test_gpu_parity import make

for name, small, B in [("LiH", True, 32), ("LiH", False, 32), ("N2", False, 24), ("HChain10", False, 8)]:
    phys, d, p32, p64, R, r, eng = make(name, B, small=small)
    ref = om.forward_laplacian(p64, d, r.double(), R.double(), phys.Z)
    f32 = om.forward_laplacian(p32, d, r, R, phys.Z)
    e_loc, aux = eng.local_energy(r.cuda(), with_aux=True)
    lp = eng.log_psi_sqr(r.cuda())[1].double().cpu()
    e_loc = e_loc.double().cpu()
    rl = (lp - ref["logpsi2"]).abs() / ref["logpsi2"].abs()
    rl32 = (f32["logpsi2"].double() - ref["logpsi2"]).abs() / ref["logpsi2"].abs()
    sc = ref["E_loc"].abs().clamp_min(1.0)
    er = (e_loc - ref["E_loc"]).abs() / sc
    er32 = (f32["E_loc"].double() - ref["E_loc"]).abs() / sc
    q = lambda t: [float(t.median()), float(t.quantile(0.9)), float(t.max())]
    print(name, small, B, "\n  logpsi2 gpu med/p90/max", q(rl), " fp32cpu", q(rl32), "\n  E_loc gpu", q(er), " fp32cpu", q(er32),
          "\n  ratio gpu/max(1e-4,3*floor) max", float((er / torch.maximum(torch.full_like(er, 1e-4), 3 * er32)).max()),
          " lp ratio", float((rl / torch.maximum(torch.full_like(rl, 1e-5), 3 * rl32)).max()))
