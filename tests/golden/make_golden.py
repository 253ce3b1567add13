"""Regenerates the committed fixtures under tests/golden/.

The reference (jax 0.4.23 / haiku / folx) cannot be imported in this image, so these are NOT outputs of the
reference: threefry_kat.json holds published known answers (Random123 threefry2x32-20 KATs and upstream-jax
facts); the .npz files are fp64 outputs of oracle/ and pin it (and the CUDA path) against
regressions.  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import model as om, mcmc as omc, threefry  # noqa: E402

OUT = Path(__file__).resolve().parent

KAT = {
    "threefry2x32": [
        {"key": [0, 0], "ctr": [0, 0], "out": [0x6B200159, 0x99BA4EFE]},
        {"key": [0xFFFFFFFF, 0xFFFFFFFF], "ctr": [0xFFFFFFFF, 0xFFFFFFFF], "out": [0x1CB996FC, 0xBB002BE7]},
        {"key": [0x13198A2E, 0x03707344], "ctr": [0x243F6A88, 0x85A308D3], "out": [0xC4923A9C, 0x483DF7A0]},
    ],
    "jax": {
        "split_PRNGKey0": [[4146024105, 967050713], [2718843009, 1272950319]],
        "uniform_PRNGKey0": 0.41845703,
        "normal_PRNGKey0": -0.20584226,
        "normal_PRNGKey42": -0.18471177,
    },
}

CASES = {
    "LiH_small": dict(R=[[0, 0, 0], [3.015, 0, 0]], Z=[3, 1], n_up=2, mapping=[0, 1, 0, 0],
                      dims=dict(n_iterations=2, n_hidden_one_el=[16, 16], n_hidden_two_el=[4], emb_dim=8, n_dets=3)),
    "LiH": dict(R=[[0, 0, 0], [3.015, 0, 0]], Z=[3, 1], n_up=2, mapping=[0, 1, 0, 0], dims={}),
    # transferable atomic orbitals from a synthetic geometry cache (config_bm_hfcoeff.yml: 4 determinants)
    "LiH_tao": dict(R=[[0, 0, 0], [3.015, 0, 0]], Z=[3, 1], n_up=2, mapping=[0, 1, 0, 0], dims=dict(n_dets=4, use_taos=True)),
}


def make_model_fixture(name, spec, B=8):
    d = om.ModelDims(n_el=len(spec["mapping"]), n_up=spec["n_up"], n_ion=len(spec["Z"]), Z_max=max(spec["Z"]), **spec["dims"])
    params = om.cast_params(om.cast_params(om.init_params(d, seed=11, bias_scale=0.1, envelope_jitter=0.5), torch.float32), torch.float64)
    g = torch.Generator().manual_seed(5)
    R = torch.tensor(spec["R"], dtype=torch.float32)
    r = (R[torch.tensor(spec["mapping"])][None] + torch.randn(B, d.n_el, 3, generator=g)).float()
    tao = om.cast_tao_cache(om.cast_tao_cache(om.make_tao_cache(d, seed=11), torch.float32), torch.float64) if d.use_taos else None
    out = om.forward_laplacian(params, d, r.double(), R.double(), spec["Z"], tao=tao)
    np.savez_compressed(OUT / f"model_{name}.npz", r=r.numpy(), R=R.numpy(), Z=np.array(spec["Z"]), n_up=spec["n_up"],
                        seed=11, bias_scale=0.1, envelope_jitter=0.5,
                        logpsi2=out["logpsi2"].numpy(), phase=out["phase"].numpy(), grad=out["grad"].numpy(),
                        E_kin=out["E_kin"].numpy(), E_pot=out["E_pot"].numpy(), E_loc=out["E_loc"].numpy())


def make_mcmc_fixture():
    B, N = 16, 4
    keys = threefry.split(threefry.prng_key(1234), B)
    r0 = threefry.normal(threefry.prng_key(7), (B, N, 3))
    func = lambda r: (-np.sum(r.astype(np.float32) ** 2, axis=(1, 2))).astype(np.float32)   # Gaussian log-density
    st = omc.OracleMCMCState(r=r0, R=np.zeros((1, 3), np.float32), Z=np.array([4]), log_psi_sqr=-np.ones(B, np.float32) * 1000,
                             walker_age=np.zeros(B, np.int32), rng_state=keys, stepsize=np.float32(0.3))
    out = omc.run_mcmc_steps(func, st, 7, max_age=2, stepsize_update_interval=3)
    np.savez_compressed(OUT / "mcmc_gaussian.npz", keys0=keys, r0=r0, keys=out.rng_state, r=out.r, age=out.walker_age,
                        log_psi_sqr=out.log_psi_sqr, stepsize=out.stepsize, acc_rate=out.acc_rate, step_nr=out.step_nr)


if __name__ == "__main__":
    (OUT / "threefry_kat.json").write_text(json.dumps(KAT, indent=1))
    for name, spec in CASES.items():
        make_model_fixture(name, spec)
    make_mcmc_fixture()
    print("wrote", sorted(p.name for p in OUT.glob("*.*")))
