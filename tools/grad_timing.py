"""Times dpe_param_gradient (gradient only / gradient + KFAC factors) on N2 x 4096 and benzene x 1024 walkers."""
import sys, time, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_gpu_parity import make
for name, B in [("N2", 4096), ("Benzene", 1024)]:
    phys, d, p32, p64, R, r, eng = make(name, B)
    r = r.cuda(); cot = torch.randn(B, device="cuda") / B
    for kf in (False, True):
        for _ in range(2): eng.param_gradient(r, cot, with_kfac=kf)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): eng.param_gradient(r, cot, with_kfac=kf)
        e1.record(); torch.cuda.synchronize()
        print(name, B, "kfac" if kf else "grad", e0.elapsed_time(e1) / 5, "ms; workspace MB", eng._ws.numel() / 2**20)
