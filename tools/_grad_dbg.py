import sys, torch, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_gpu_parity import make
from oracle import gradient as og, model as om
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for name, small, B, seed in [("LiH", False, 16, 3), ("LiH", False, 16, 4), ("LiH", False, 64, 5), ("LiH", True, 24, 3), ("N2", False, 6, 3), ("N2", False, 12, 7)]:
    phys, d, p32, p64, R, r, eng = make(name, B, small=small, seed=seed)
    g = torch.Generator().manual_seed(5)
    cot = (torch.randn(B, generator=g) / B).float()
    flat, lp = eng.param_gradient(r.cuda(), cot.cuda(), with_kfac=True)
    ref = og.param_gradient(p64, d, r.double(), R.double(), phys.Z, cot.double())
    ref32 = og.param_gradient(p32, d, r, R, phys.Z, cot)
    E, F = [], []
    for (mod, leaf), (off, size, rows, cols) in zip(eng.leaves, eng.leaf_shapes):
        got = flat[off:off + size].reshape(ref[mod][leaf].shape)
        E.append(rel(got, ref[mod][leaf])); F.append(rel(ref32[mod][leaf], ref[mod][leaf]))
    E, F = np.array(E), np.array(F)
    lp64 = om.log_psi_sqr(p64, d, r.double(), R.double(), phys.Z)[1]
    lp32 = om.log_psi_sqr(p32, d, r, R, phys.Z)[1]
    cond = om.cond(p64, d, r.double(), R.double(), phys.Z) if hasattr(om, "cond") else None
    print(f"{name} small={small} B={B} seed={seed}: grad err med {np.median(E):.2e} max {E.max():.2e} | cpu32 med {np.median(F):.2e} max {F.max():.2e} | "
          f"lp err gpu {float((lp.cpu().double()-lp64).abs().max()):.2e} cpu32 {float((lp32.double()-lp64).abs().max()):.2e}")
    fac = og.kfac_factors(p64, d, r.double(), R.double(), phys.Z); fac32 = og.kfac_factors(p32, d, r, R, phys.Z)
    kf = flat[eng.n_params:]
    EG, FG = [], []
    for lname, din, dout, hb, rpw, a_off, g_off in eng.kfac_layers():
        G = kf[g_off:g_off + dout * dout].reshape(dout, dout)
        EG.append(rel(G, fac[lname][1])); FG.append(rel(fac32[lname][1], fac[lname][1]))
    print(f"     G err med {np.median(EG):.2e} max {max(EG):.2e} | cpu32 med {np.median(FG):.2e} max {max(FG):.2e}")
