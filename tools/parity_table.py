"""Error-vs-conditioning table of the CUDA path against the fp64 oracle (profiles/rNN_parity_table.md): per cond_eff decade the
median / max relative error of the tcgen05 3xTF32 path, the FP32 SIMT path and the fp32 CPU restatement (the fp32 floor).
Run on the GPU box:  python tools/parity_table.py gpurun_out/parity_table.md"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np
import torch
from oracle import model as om, parity_rule
from test_gpu_parity import make

out = [
    "# CUDA path vs fp64 oracle by conditioning of the walker",
    "",
    "cond_eff = sum_d |q_d| cond_2(A_d) / |sum_d q_d| (oracle/model.py). Errors are relative (E_loc: |dE| / max(|E|, 1)).",
    "`tc` = tcgen05 3xTF32 dense layers + tensor-core determinant traces (default), `simt` = FP32 CUDA-core GEMMs and determinant stage,",
    "`fp32 cpu` = the oracle's algorithm in fp32 on the CPU (what any fp32 evaluation loses at that conditioning).",
    "",
    "| case | quantity | cond_eff | walkers | tc median | tc max | simt median | simt max | fp32 cpu median | fp32 cpu max |",
    "|---|---|---|---:|---:|---:|---:|---:|---:|---:|",
]
worst = []


def rows(case, what, e_tc, e_simt, e_f32, cond):
    dec = np.ceil(np.log10(np.maximum(cond, 1.0))).astype(int)
    for d in sorted(set(dec.tolist())):
        m = dec == d
        out.append(f"| {case} | {what} | <1e{d} | {int(m.sum())} | {np.median(e_tc[m]):.1e} | {e_tc[m].max():.1e} | {np.median(e_simt[m]):.1e} | "
                   f"{e_simt[m].max():.1e} | {np.median(e_f32[m]):.1e} | {e_f32[m].max():.1e} |")


def run(case, phys, d, p32, p64, R, r, eng):
    ref = om.forward_laplacian(p64, d, r.double(), R.double(), phys.Z)
    f32 = om.forward_laplacian(p32, d, r, R, phys.Z)
    cond = ref["cond"].numpy()
    res = {}
    for tag, gp, simt in (("tc", 1, False), ("simt", 0, True)):
        eng.set_gemm_path(gp)
        eng.set_det_path(simt=simt)
        e = eng.local_energy(r.cuda()).double().cpu()
        lp = eng.log_psi_sqr(r.cuda())[1].double().cpu()
        res[tag] = (((lp - ref["logpsi2"]).abs() / ref["logpsi2"].abs()).numpy(), ((e - ref["E_loc"]).abs() / ref["E_loc"].abs().clamp_min(1.0)).numpy())
    fl = (((f32["logpsi2"].double() - ref["logpsi2"]).abs() / ref["logpsi2"].abs()).numpy(),
          ((f32["E_loc"].double() - ref["E_loc"]).abs() / ref["E_loc"].abs().clamp_min(1.0)).numpy())
    rows(case, "log psi^2", res["tc"][0], res["simt"][0], fl[0], cond)
    rows(case, "E_loc", res["tc"][1], res["simt"][1], fl[1], cond)
    for k, tol in ((0, 1e-5), (1, 1e-4)):
        b = parity_rule.bounds(fl[k], cond, tol)
        worst.append((case, "log psi^2" if k == 0 else "E_loc", float((res["tc"][k] / b).max()), float((res["simt"][k] / b).max()),
                      float(np.quantile(res["tc"][k], 0.99)), float(np.quantile(fl[k], 0.99))))


for name, B in (("LiH", 96), ("N2", 96), ("HChain10", 24), ("Allene_TinyMol", 12), ("Benzene", 6)):
    phys, d, p32, p64, R, r, eng = make(name, B)
    run(f"{name}, {B} Gaussian walkers", phys, d, p32, p64, R, r, eng)

# |psi|^2-distributed walkers (1000 Metropolis steps)
import deeperwin_b200 as dpe
for name, B in (("N2", 128), ("LiH", 128)):
    cfg = dpe.Configuration(physical=dict(name=name))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=11, device="cuda:0")
    st = dpe.MCMCState.initialize_around_nuclei(B, phys, "exponential", "el_ion_mapping", dpe.PRNGKey(5), device="cuda:0")
    st = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=1000)).run_inter_steps(f, st, params, phys.n_up, phys.n_dn, fixed)
    d = om.ModelDims(n_el=phys.n_electrons, n_up=phys.n_up, n_ion=len(phys.Z), Z_max=max(phys.Z))
    p32 = {m: {k: v.cpu() for k, v in l.items()} for m, l in params.items()}
    f.engine.set_params(params); f.engine.set_geometry(st.R, st.Z)
    run(f"{name}, {B} walkers after 1000 Metropolis steps", phys, d, p32, om.cast_params(p32, torch.float64), st.R.cpu(), st.r.cpu(), f.engine)

out += ["", "## Worst ratio error / parity-rule bound (must be <= 1) and 99th percentiles", "",
        "| case | quantity | tc max(err/bound) | simt max(err/bound) | tc p99 | fp32 cpu p99 |", "|---|---|---:|---:|---:|---:|"]
out += [f"| {c} | {w} | {a:.2f} | {b:.2f} | {p:.1e} | {q:.1e} |" for c, w, a, b, p, q in worst]
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 1:
    Path(sys.argv[1]).parent.mkdir(parents=True, exist_ok=True)
    Path(sys.argv[1]).write_text(text)
