"""Error-vs-conditioning table of the CUDA path against the fp64 oracle (profiles/rNN_parity_table.md): per cond_eff decade the
median / max relative error of the tcgen05 3xTF32 path, the FP32 SIMT path, one fp32 CPU evaluation and the per-walker fp32
floor (envelope over same-spin electron permutations, oracle/parity_rule.py), then the parity rule's verdict per case.
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
    "`fp32 cpu` = one evaluation of the oracle's algorithm in fp32 on the CPU, `floor` = the per-walker fp32 floor of the parity rule",
    f"(largest error over {parity_rule.N_PERM} same-spin electron permutations of that fp32 CPU evaluation).",
    "",
    "| case | quantity | cond_eff | walkers | tc median | tc max | simt median | simt max | fp32 cpu median | fp32 cpu max | floor median | floor max |",
    "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|",
]
verdict = []


def run(case, phys, d, p32, p64, R, r, eng):
    ref = om.forward_laplacian(p64, d, r.double(), R.double(), phys.Z)
    env = parity_rule.fp32_envelope(om, p32, d, r, R, phys.Z, ref)
    one = parity_rule.errors(om.forward_laplacian(p32, d, r, R, phys.Z), ref)
    cond = ref["cond"].numpy()
    res = {}
    for tag, gp, simt in (("tc", 1, False), ("simt", 0, True)):
        eng.set_gemm_path(gp)
        eng.set_det_path(simt=simt)
        e = eng.local_energy(r.cuda())
        lp = eng.log_psi_sqr(r.cuda())[1]
        res[tag] = parity_rule.errors(dict(logpsi2=lp, E_loc=e), ref)
    for q, label, tol in (("logpsi2", "log psi^2", 1e-5), ("E_loc", "E_loc", 1e-4)):
        out.extend(parity_rule.table_rows(case, label, cond, [res["tc"][q], res["simt"][q], one[q], env[q]]))
        row = [case, label]
        for tag in ("tc", "simt"):
            try:
                n = parity_rule.check(res[tag][q], env[q], tol, cond=ref["cond"] if q == "logpsi2" else None)
                row.append(f"pass ({n} at plain tol)")
            except AssertionError as ex:
                row.append("FAIL " + str(ex)[:160].replace("|", "/"))
            hard = np.maximum(tol, parity_rule.HARD_FACTOR * env[q])
            row.append(f"{(res[tag][q] / hard).max():.2f} / {max(np.quantile(res[tag][q], x) / max(tol, parity_rule.FLOOR_FACTOR * np.quantile(env[q], x)) for x in parity_rule.QUANTILES):.2f}")
        row += [f"{np.quantile(res['tc'][q], 0.99):.1e}", f"{np.quantile(env[q], 0.99):.1e}"]
        verdict.append(row)


for name, B in (("LiH", 96), ("N2", 96), ("HChain10", 24), ("Allene_TinyMol", 12), ("Benzene", 6)):
    phys, d, p32, p64, R, r, eng = make(name, B)
    run(f"{name}, {B} Gaussian walkers", phys, d, p32, p64, R, r, eng)
    del eng

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

out += ["", "## Parity rule per case (oracle/parity_rule.py) and 99th percentiles", "",
        "Ratios (must be <= 1): worst walker err / max(tol, 16 x its floor)  /  worst of the q = 0.5, 0.9, 0.99 quantile ratios err_q / max(tol, 2 x floor_q).", "",
        "| case | quantity | tc rule | tc ratios | simt rule | simt ratios | tc p99 | floor p99 |", "|---|---|---|---:|---|---:|---:|---:|"]
out += ["| " + " | ".join(r) + " |" for r in verdict]
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 1:
    Path(sys.argv[1]).parent.mkdir(parents=True, exist_ok=True)
    Path(sys.argv[1]).write_text(text)
