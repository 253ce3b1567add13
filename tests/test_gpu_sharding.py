"""Sharded execution equals single-device execution.  (1) On one GPU: the walkers split into two shards that run through
dpe_mcmc_steps(run_controller = 0), their accept counts summed and dpe_mcmc_controller replayed -- exactly what two ranks do
(deeperwin_b200/mcmc.py, plan_segments) -- reproduce the full-batch chain with the in-call controller bit for bit.
(2) With two GPUs and NCCL (skipped on one): MetropolisHastingsMonteCarlo + build_total_energy on sharded walkers vs the
single-process result."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_two_shards_on_one_gpu_equal_the_full_batch_chain():
    import deeperwin_b200 as dpe
    from deeperwin_b200._lib import DpeMcmcState
    from deeperwin_b200.mcmc import plan_segments
    cfg = dpe.Configuration(physical=dict(name="LiH"))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=5, device="cuda:0")
    eng = f.engine
    B, n_steps, interval = 128, 23, 5
    st0 = dpe.MCMCState.initialize_around_nuclei(B, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(3), device="cuda:0")
    st0.stepsize = torch.tensor(0.2, device="cuda")
    mcfg = dpe.MCMCConfigOptimization(n_inter_steps=n_steps, stepsize_update_interval=interval, max_age=4, initialization="gaussian")
    mc = dpe.MetropolisHastingsMonteCarlo(mcfg)
    full = mc.run_inter_steps(f, st0, params, 2, 2, fixed)                  # one device: in-call controller

    # two shards, driven as two ranks drive them
    eng.set_params(params); eng.set_geometry(st0.R, st0.Z)
    half = B // 2
    shards = []
    for k in range(2):
        sl = slice(k * half, (k + 1) * half)
        t = dict(r=st0.r[sl].clone(), lp=st0.log_psi_sqr[sl].clone(), age=st0.walker_age[sl].clone(), keys=st0.rng_state[sl].clone(),
                 ss=st0.stepsize.reshape(1).clone(), sn=st0.step_nr.reshape(1).clone(), ar=st0.acc_rate.reshape(1).clone())
        t["st"] = DpeMcmcState(t["r"].data_ptr(), t["lp"].data_ptr(), t["age"].data_ptr(), t["keys"].data_ptr(), t["ss"].data_ptr(), t["sn"].data_ptr(), t["ar"].data_ptr())
        shards.append(t)
    done, all_counts = 0, []
    for seg in plan_segments(0, n_steps, interval):
        counts = [torch.zeros(seg, dtype=torch.int32, device="cuda") for _ in range(2)]
        for t, c in zip(shards, counts):
            eng.mcmc_steps(t["st"], half, seg, mc._cfg, done == 0, False, c)
        total = counts[0] + counts[1]                                         # the all-reduce of two ranks
        for t in shards:
            eng.mcmc_controller(t["st"], total, seg, B, mc._cfg)
        all_counts.append(total)
        done += seg
    torch.cuda.synchronize()
    for key, field in (("r", "r"), ("lp", "log_psi_sqr"), ("age", "walker_age")):
        assert torch.equal(torch.cat([shards[0][key], shards[1][key]]), getattr(full, field)), field
    assert torch.equal(torch.cat([shards[0]["keys"], shards[1]["keys"]]).view(torch.int32), full.rng_state.view(torch.int32))
    for t in shards:                                                          # replicated scalars agree with the single-device chain
        assert t["ss"].item() == full.stepsize.item() and int(t["sn"]) == int(full.step_nr) == n_steps and t["ar"].item() == full.acc_rate.item()
    assert torch.equal(torch.cat(all_counts), mc.last_accept_counts)


def _nccl_worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        import deeperwin_b200 as dpe
        dev = torch.device(f"cuda:{rank}")
        cfg = dpe.Configuration(physical=dict(name="LiH"))
        phys = cfg.physical
        f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=5, device=dev)
        full = dpe.MCMCState.initialize_around_nuclei(128, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(3), device=dev)
        mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=12, stepsize_update_interval=5, initialization="gaussian"))
        st = mc.run_inter_steps(f, full.split_across_devices(), params, 2, 2, fixed)
        te = dpe.build_total_energy(dpe.build_local_energy(f, forward_lap=True), dpe.ClippingConfig(center="median", width_metric="mae", clip_by=3.0))
        loss, (cs, aux) = te(params, dpe.init_clipping_state(device=dev), (2, 2), (st.r[0], st.R[0], st.Z[0], fixed))
        merged = st.merge_devices()
        ret[rank] = dict(r=merged.r.cpu(), age=merged.walker_age.cpu(), stepsize=merged.stepsize.item(), acc=merged.acc_rate.item(),
                         E_mean=aux["E_mean"].item(), E_var=aux["E_var"].item(), Ec=aux["E_mean_clipped"].item(), center=cs[0].item(), width=cs[1].item(),
                         E_loc=aux["E_loc"].cpu())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_two_ranks_equal_single_process():
    import torch.multiprocessing as mp
    import deeperwin_b200 as dpe
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, 29600 + os.getpid() % 1000, ret), nprocs=2, join=True)
    cfg = dpe.Configuration(physical=dict(name="LiH"))
    phys = cfg.physical
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=5, device="cuda:0")
    full = dpe.MCMCState.initialize_around_nuclei(128, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(3), device="cuda:0")
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=12, stepsize_update_interval=5, initialization="gaussian"))
    st = mc.run_inter_steps(f, full, params, 2, 2, fixed)
    e = dpe.build_local_energy(f, forward_lap=True)(params, (2, 2), st.r, st.R, st.Z, fixed).cpu()
    for rank in range(2):
        o = ret[rank]
        assert torch.equal(o["r"], st.r.cpu()) and torch.equal(o["age"], st.walker_age.cpu())
        assert o["stepsize"] == st.stepsize.item() and abs(o["acc"] - st.acc_rate.item()) < 1e-7
        assert torch.equal(o["E_loc"], e[rank * 64:(rank + 1) * 64])
        assert abs(o["E_mean"] - e.mean().item()) < 1e-5 * abs(e.mean().item())
        # median centre: pmean of the per-rank medians (loss_function.py:20-21), not the median of the union
        assert np.isfinite([o["E_var"], o["Ec"], o["center"], o["width"]]).all()
