"""world_size-2 gloo tests of the N>1 host logic (walker sharding, the natural reductions)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeperwin_b200 import utils
        from deeperwin_b200.mcmc import MCMCState
        from oracle import mcmc as omc
        out = {}
        x = torch.tensor([1.0 + rank, 10.0 * (rank + 1)])
        out["pmean"] = utils.pmean(x).tolist()
        out["psum"] = utils.psum(x).tolist()
        assert x.tolist() == [1.0 + rank, 10.0 * (rank + 1)]          # functional: input untouched
        B, N = 8, 3
        g = torch.Generator().manual_seed(0)
        full = MCMCState(r=torch.randn(B, N, 3, generator=g), R=torch.zeros(2, 3), Z=torch.tensor([1, 1], dtype=torch.int32),
                         log_psi_sqr=torch.arange(B, dtype=torch.float32), walker_age=torch.arange(B, dtype=torch.int32),
                         rng_state=torch.arange(2 * B, dtype=torch.int32).view(torch.uint32).reshape(B, 2))
        sp = full.split_across_devices()
        assert sp.r.shape == (1, B // world, N, 3) and sp.R.shape == (1, 2, 3) and sp.stepsize.shape == (1,)
        assert torch.equal(sp.r[0], full.r[rank * 4:(rank + 1) * 4])      # contiguous blocks (mcmc.py:105-129)
        merged = sp.merge_devices()
        for k in ("r", "log_psi_sqr", "walker_age"):
            assert torch.equal(getattr(merged, k), getattr(full, k)), k
        assert torch.equal(merged.rng_state.view(torch.int32), full.rng_state.view(torch.int32))
        # flat all-reduce of "gradients"
        ts = [torch.full((3,), float(rank)), torch.full((2, 2), 2.0 * rank)]
        utils.flat_allreduce_mean(ts)
        out["flat"] = [t.flatten().tolist() for t in ts]
        # energy statistics with pmean == statistics of the equal-sized shards combined as the reference does
        rng = np.random.default_rng(1)
        E = rng.normal(-5, 2, 64).astype(np.float32)
        shard = E[rank * 32:(rank + 1) * 32]
        ar = lambda v: float(utils.pmean(torch.tensor([float(v)], dtype=torch.float32))[0])
        loss, st, aux = omc.energy_statistics(shard, omc.init_clipping_state(), allreduce_mean=ar)
        out["E_mean"], out["E_var"] = float(aux["E_mean"]), float(aux["E_var"])
        out["E_ref_mean"], out["E_ref_var"] = float(E.mean()), float(E.var())
        # median-centred / MAE window (the sample configs): centre = all-reduced mean of the per-rank medians, exactly the
        # reference's pmean(nanmedian(E)) (loss_function.py:20-21) -- not the median of the union
        _, st_med, _ = omc.energy_statistics(shard, (np.float32(-5.0), np.float32(6.0)), center="median", width_metric="mae",
                                             clip_by=3.0, allreduce_mean=ar)
        out["med_center"], out["med_width"] = float(st_med[0]), float(st_med[1])
        clipped = [(-5.0 + np.tanh((E[k * 32:(k + 1) * 32] + 5.0) / 6.0) * 6.0).astype(np.float32) for k in range(world)]
        c_ref = np.float32(np.mean([np.float32(np.median(c)) for c in clipped]))
        out["med_center_ref"] = float(c_ref)
        out["med_width_ref"] = float(3.0 * np.mean([np.mean(np.abs(c - c_ref)) for c in clipped]))
        # deferred accept-count all-reduce (mcmc.plan_segments): each rank runs its shard of an oracle chain without a controller,
        # the counts are all-reduced per segment and the scalar tail of make_mcmc_step (mcmc.py:367-377) is replayed from them;
        # the result must equal the chain on the union of the walkers with the per-step pmean of the reference
        from deeperwin_b200.mcmc import plan_segments
        from oracle import threefry
        Bt, Ne, n_steps, interval = 16, 2, 13, 4
        keys = threefry.split(threefry.prng_key(5), Bt)
        r0 = rng.normal(0, 1, (Bt, Ne, 3)).astype(np.float32)
        func = lambda rr: (-2.0 * np.linalg.norm(rr, axis=-1).sum(-1)).astype(np.float32)          # log psi^2 of a product of 1s orbitals
        mk = lambda sl: omc.OracleMCMCState(r=r0[sl].copy(), R=np.zeros((1, 3), np.float32), Z=np.array([2]), log_psi_sqr=func(r0[sl]),
                                            walker_age=np.zeros(len(r0[sl]), np.int32), rng_state=keys[sl].copy(), stepsize=np.float32(0.5))
        union = mk(slice(None))
        for _ in range(n_steps):
            union = omc.make_mcmc_step(func, union, max_age=3, stepsize_update_interval=interval)
        shard = mk(slice(rank * 8, (rank + 1) * 8))
        ss, ar, sn = np.float32(0.5), np.float32(0.0), 0
        for seg in plan_segments(0, n_steps, interval):
            counts = []
            for _ in range(seg):
                shard.stepsize = ss                                              # replicated scalar: changes only at segment boundaries
                shard, mask = omc.make_mcmc_step(func, shard, max_age=3, stepsize_update_interval=10 ** 9, return_mask=True)
                counts.append(int(mask.sum()))
            tot = utils.psum(torch.tensor(counts, dtype=torch.int32)).numpy()
            for c in tot:                                                        # dpe_mcmc_controller (csrc/mcmc.cu k_controller)
                rate = np.float32(c) / np.float32(Bt)
                sn += 1
                ar_new = np.float32(np.float32(0.9) * ar + np.float32(0.1) * rate)
                if sn % interval == 0:
                    ss = np.float32(ss / np.float32(1.05)) if ar < np.float32(0.5) else np.float32(ss * np.float32(1.05))
                    ss = np.float32(np.clip(ss, np.float32(0.01), np.float32(1.0)))
                ar = ar_new
        out["chain_r_equal"] = bool(np.array_equal(shard.r, union.r[rank * 8:(rank + 1) * 8]))
        out["chain_scalars"] = (float(ss), float(ar), sn, float(union.stepsize), float(union.acc_rate), union.step_nr)
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        o = ret[rank]
        assert o["pmean"] == [1.5, 15.0] and o["psum"] == [3.0, 30.0]
        assert o["flat"] == [[0.5] * 3, [1.0] * 4]
        assert abs(o["E_mean"] - o["E_ref_mean"]) < 1e-5 and abs(o["E_var"] - o["E_ref_var"]) < 1e-4
        assert abs(o["med_center"] - o["med_center_ref"]) < 1e-5 and abs(o["med_width"] - o["med_width_ref"]) < 1e-4
        assert o["chain_r_equal"]
        ss, ar, sn, uss, uar, usn = o["chain_scalars"]
        assert ss == uss and sn == usn and abs(ar - uar) < 1e-7
