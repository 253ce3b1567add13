#!/usr/bin/env python
"""bench.py -- walker E_loc evaluations / second of the VMC inner loop (BASELINE.json metric).

One *step* = one pass of the hot path over one batch of synthetic walkers: one Metropolis-Hastings step
(mcmc.py:345-379: threefry proposal + forward pass + accept/reject) followed by one local-energy evaluation
(hamiltonian.py:272-291, forward Laplacian) and the E_loc statistics (loss_function.py:89-109) -- i.e. one
"eval" per walker.  Workload at N=1: BASELINE.json configs[1] = N2, 4096 walkers, default dpe4
model, random-init weights, walkers burnt in for 100 steps; N>1 keeps 4096 walkers per GPU (weak scaling,
walkers are independent chains; the only collectives are the natural scalar reductions).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--molecule N2] [--walkers 4096]
  python bench.py --impl reference ...   # the CPU restatement of the reference (oracle/) on the host cores

`value`   : device-timed throughput, inputs resident in HBM (CUDA events around every step, L2 flushed between steps)
`e2e`     : same metric through the public Python API (MetropolisHastingsMonteCarlo.run_inter_steps + get_local_energy)
            with HOST buffers: pinned H2D of the walker state and D2H of the new state + E_loc inside the timed region
`roofline`: the dominant kernel (the dense-layer GEMM), timed live with CUDA events inside the library
`cpu_baseline`: oracle/ (a port, not the reference binary: jax is not installable here) on a bounded sample
`secondary`: the other half of BASELINE.json's metric, Benzene (42 electrons) x 4096 walkers per GPU (configs[3]), fewer steps
`cadence`  : the reference's optimisation cadence -- n_inter_steps = 20 Metropolis steps per E_loc evaluation (configuration.py:1039)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# algorithmic FLOP per walker-eval (SURVEY.md 8d / BASELINE.md section 2)
FLOP_PER_EVAL = {"LiH": 0.0872e9, "HChain10": 0.533e9, "N2": 1.027e9, "Benzene": 12.20e9}
METRIC = "walker_eloc_evals_per_sec"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--molecule", default="N2")
    ap.add_argument("--walkers", type=int, default=4096, help="walkers per GPU")
    ap.add_argument("--burn-in", type=int, default=100)
    ap.add_argument("--gemm-path", type=int, default=int(os.environ.get("DPE_GEMM_PATH", "-1")),
                    help="-1 library default, 0 FP32 SIMT, 1 tcgen05 3xTF32")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--secondary", default="Benzene", help="molecule of the secondary block ('' to skip)")
    ap.add_argument("--secondary-steps", type=int, default=3)
    ap.add_argument("--no-cadence", action="store_true")
    ap.add_argument("--no-weight-sharing", action="store_true")
    return ap.parse_args()


def workload_string(molecule, n_el, walkers):
    return (f"{molecule} ({n_el} electrons), {walkers} walkers per GPU, default dpe4 model (4x256/32, 32 full determinants), "
            "random-init weights, 100 burn-in steps; step = 1 Metropolis step + 1 forward-Laplacian E_loc + E statistics")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm_gbs=d["hbm_gbs"], sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


# ------------------------------------------------------------------------------------------- CPU arm
def _cpu_port(molecule: str):
    import torch
    from oracle import model as om
    from deeperwin_b200.configuration import PhysicalConfig
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    phys = PhysicalConfig(name=molecule)
    d = om.ModelDims(n_el=phys.n_electrons, n_up=phys.n_up, n_ion=len(phys.Z), Z_max=max(phys.Z))
    p32 = om.init_params(d, seed=1234, dtype=torch.float32)
    R = torch.tensor(phys.R, dtype=torch.float32)
    g = torch.Generator().manual_seed(1234)

    def run(n):
        """n walkers: one forward pass (the Metropolis step's log psi^2) + one forward-Laplacian E_loc in chunks of 64 (hamiltonian.py:280)."""
        r = (R[torch.tensor(phys.el_ion_mapping)][None] + torch.randn(n, d.n_el, 3, generator=g)).float()
        t0 = time.perf_counter()
        with torch.no_grad():
            om.log_psi_sqr(p32, d, r, R, phys.Z)
            om.local_energy(p32, d, r, R, phys.Z, max_batch_size=64)
        return time.perf_counter() - t0

    return run, cores, phys


def cpu_port_throughput(molecule: str, target_seconds: float):
    """Times oracle/ (torch CPU fp32, all host threads) on a bounded sample of the workload."""
    run, cores, _ = _cpu_port(molecule)
    run(16)                       # warm-up (thread pools, allocator)
    pilot = run(64)
    n = int(max(64, min(8192, 64 * target_seconds / max(pilot, 1e-3)) // 64 * 64))
    t = run(n)
    return dict(value=n / t, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} {molecule} walkers (1 forward + 1 forward-Laplacian E_loc each), torch-CPU fp32 oracle, {t:.1f} s")


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  jax/haiku/folx are not installable in this image (no
    network, not in /opt/wheelhouse), so this arm times the oracle port (kind = "port": same algorithm, torch-CPU fp32, all host
    threads).  One step = the hot path over a bounded SAMPLE of the workload's walkers (the full 4096 would take minutes per
    step); `ms_per_step` is the measured time of that sample, `value` = sample walkers / that time.  With --gpus N > 1 the arm
    is still ONE host (rank 0 runs, the other ranks exit): only the N = 1 ratio compares like with like."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, cores, phys = _cpu_port(args.molecule)
    run(16)
    pilot = run(64)
    n = int(max(64, min(4096, 64 * min(args.cpu_seconds, 6.0) / max(pilot, 1e-3)) // 64 * 64))
    times = [run(n) for _ in range(args.warmup + args.steps)][args.warmup:]
    t = sum(times) / len(times)
    v = n / t
    info = dict(value=v, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} of the {args.walkers} {args.molecule} walkers per step (1 forward + 1 forward-Laplacian E_loc each), torch-CPU fp32 oracle, "
                       f"{t:.1f} s per step; one host regardless of --gpus")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.molecule, phys.n_electrons, args.walkers), "walkers_per_gpu": args.walkers,
                       "sample_walkers_per_step": n},
            "cpu_baseline": info, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML is queried in-process from a background thread
    (pynvml): spawning `nvidia-smi -lms` was measured to stall the GPU for 60-100 ms per start-up/poll, which lands
    inside 40 ms steps; the in-process queries do not.  Falls back to one-shot nvidia-smi samples if pynvml is missing."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index, period=float(os.environ.get("DPE_BENCH_CLOCK_PERIOD", "0.05"))):
        self.rows, self.stop_flag, self.h = [], False, None
        self.period = period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:                                   # the CUDA device this rank uses, whatever CUDA_VISIBLE_DEVICES maps it to
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        while not self.stop_flag:
            t = time.perf_counter()
            try:
                if self.nv:
                    clk = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    try:
                        mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.rows.append((t, clk, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def window(self, t0, t1):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["pynvml unavailable"]}
        sel = [(c, m) for t, c, m in list(self.rows) if t0 <= t <= t1]
        sm = sorted(c for c, _ in sel)
        reasons = sorted(n for n, bit in self.REASONS.items() if any(m & bit for _, m in sel))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm)}

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=1.0)


class Workload:
    """One molecule x walkers-per-GPU on this rank: model, burnt-in walker state resident in HBM, and the timed step."""

    def __init__(self, args, molecule, walkers, burn_in, world, rank, dev):
        import torch
        import deeperwin_b200 as dpe
        from deeperwin_b200._lib import DpeMcmcState
        self.torch, self.dpe, self.world, self.dev, self.B = torch, dpe, world, dev, walkers
        cfg = dpe.Configuration(physical=dict(name=molecule), optimization=dict(mcmc=dict(n_walkers=walkers * world, initialization="gaussian")))
        self.phys = phys = cfg.physical
        self.spin = (phys.n_up, phys.n_dn)
        self.log_psi_sqr, _, _, self.params, self.fixed = dpe.build_log_psi_squared(cfg.model, phys, None, None, rng_seed=1234, device=dev)
        self.engine = self.log_psi_sqr.engine
        if args.gemm_path >= 0:
            self.engine.set_gemm_path(args.gemm_path)
        self.get_local_energy = dpe.build_local_energy(self.log_psi_sqr, forward_lap=True)
        # N > 1: the E statistics and their two scalar NCCL all-reduces run on a side stream (they are rendezvous points of all ranks;
        # the next Metropolis step proceeds underneath), and the accept counts are all-reduced only where the step-size controller
        # needs them (every stepsize_update_interval steps, mcmc.plan_segments) instead of after every step
        self.stats_stream = torch.cuda.Stream(device=dev) if world > 1 else None
        self.total_energy = dpe.build_total_energy(self.get_local_energy, cfg.optimization.clipping, stats_stream=self.stats_stream)
        # synthetic walkers: r0 = R[el_ion_mapping] + N(0,1) (mcmc.py:64-67), threefry seed 1234, then burn-in
        full = dpe.MCMCState.initialize_around_nuclei(walkers * world, phys, "gaussian", "el_ion_mapping", dpe.PRNGKey(1234), device=dev)
        state = full.split_across_devices()
        del full
        burn = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=burn_in, initialization="gaussian"))
        self.state = state = burn.run_inter_steps(self.log_psi_sqr, state, self.params, *self.spin, self.fixed)
        self.mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=1, initialization="gaussian"))
        self.clip_state = dpe.init_clipping_state(device=dev)
        # device-resident walker state driven through the C ABI
        self.r = state.r[0].clone(); self.lp = state.log_psi_sqr[0].clone(); self.age = state.walker_age[0].clone()
        self.keys = state.rng_state[0].clone()
        self.ss = state.stepsize.reshape(1).clone(); self.sn = state.step_nr.reshape(1).to(torch.int32).clone()
        self.ar = state.acc_rate.reshape(1).clone()
        self.st = DpeMcmcState(self.r.data_ptr(), self.lp.data_ptr(), self.age.data_ptr(), self.keys.data_ptr(), self.ss.data_ptr(),
                               self.sn.data_ptr(), self.ar.data_ptr())
        self.counts = torch.zeros(32, dtype=torch.int32, device=dev)
        self.pending = torch.zeros(4096, dtype=torch.int32, device=dev)      # accept counts not yet seen by the controller (N > 1)
        self.n_pending = 0
        self.step_host = int(self.sn.item())
        self.R_, self.Z_ = state.R[0], state.Z[0]
        self.aux = None
        torch.cuda.synchronize()

    def flush_counts(self):
        """All-reduces the accept counts gathered since the last step-size boundary and replays the controller over them
        (mcmc.py:367-377 from integer counts: the same acc_rate EMA and step size as a per-step pmean)."""
        import torch.distributed as dist
        if self.n_pending:
            seg = self.pending[:self.n_pending]
            dist.all_reduce(seg)
            self.engine.mcmc_controller(self.st, seg, self.n_pending, self.B * self.world, self.mc._cfg)
            self.n_pending = 0

    def device_step(self, n_mcmc=1):
        """n_mcmc Metropolis steps + one forward-Laplacian E_loc + statistics, state resident in HBM."""
        if self.world == 1:
            self.engine.mcmc_steps(self.st, self.B, n_mcmc, self.mc._cfg, False, True, self.counts[:n_mcmc])
        else:
            from deeperwin_b200.mcmc import plan_segments
            for seg in plan_segments(self.step_host, n_mcmc, self.mc._cfg.stepsize_update_interval):
                if self.n_pending + seg > self.pending.numel():
                    self.flush_counts()
                self.engine.mcmc_steps(self.st, self.B, seg, self.mc._cfg, False, False, self.pending[self.n_pending:self.n_pending + seg])
                self.n_pending += seg
                self.step_host += seg
                if self.step_host % self.mc._cfg.stepsize_update_interval == 0:
                    self.flush_counts()
        loss, (self.clip_state, self.aux) = self.total_energy(self.params, self.clip_state, self.spin, (self.r, self.R_, self.Z_, self.fixed))
        return loss

    def optimisation_epoch(self, n_mcmc=20):
        """One epoch of the reference's optimisation loop without the optimiser's control flow: n_mcmc Metropolis steps, E_loc + clipped
        statistics, then the parameter gradient and the KFAC factors of every dense layer in one backward pass (+ one flat all-reduce)."""
        import torch.distributed as dist
        self.device_step(n_mcmc)
        diff = self.aux["E_loc_clipped"] - self.aux["E_mean_clipped"]
        cot = self.torch.nan_to_num(diff, nan=0.0) / diff.numel()
        if getattr(self, "grad_flat", None) is None:
            self.grad_flat = self.torch.empty(self.engine.n_params + int(self.engine.lib.dpe_kfac_floats(self.engine.handle)),
                                              dtype=self.torch.float32, device=self.dev)
        self.engine.param_gradient(self.r, cot, with_kfac=True, out=self.grad_flat)
        if self.world > 1:
            dist.all_reduce(self.grad_flat)
            self.grad_flat /= self.world
        return self.grad_flat

    def timed_epochs(self, reps, n_mcmc, flush):
        torch = self.torch
        self.optimisation_epoch(n_mcmc)
        torch.cuda.synchronize()
        launches0 = self.engine.launch_count()
        ms = []
        for _ in range(reps):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); self.optimisation_epoch(n_mcmc)
            if self.world > 1:
                self.flush_counts()
                torch.cuda.current_stream().wait_stream(self.stats_stream)
            a1.record()
            torch.cuda.synchronize()
            ms.append(a0.elapsed_time(a1))
        t = torch.tensor([sum(ms)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / reps, (self.engine.launch_count() - launches0) // reps

    def timed(self, steps, warmup, flush, n_mcmc=1, settle=True):
        """W warm-up steps (+ a bounded settle loop), then exactly `steps` steps timed with CUDA events, L2 flushed (untimed)
        between them; returns (device ms summed over the steps = max over ranks, per-step ms of this rank, launches, wall window)."""
        torch = self.torch
        import torch.distributed as dist
        for _ in range(warmup):
            self.device_step(n_mcmc)
        torch.cuda.synchronize()
        # settle: a fresh process sees 20-100 % slower steps for its first ~0.5 s on these boxes (power / clock ramp);
        # keep warming up (bounded) until three consecutive steps agree with the fastest one seen to 3 %
        extra, best = [], float("inf")
        for _ in range(24 if settle else 0):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); self.device_step(n_mcmc); a1.record()
            torch.cuda.synchronize()
            extra.append(a0.elapsed_time(a1)); best = min(best, extra[-1])
            settled = len(extra) >= 3 and all(t <= 1.03 * best for t in extra[-3:])
            if self.world > 1:                      # every rank must run the same number of steps (the steps contain collectives)
                flag = torch.tensor([0 if settled else 1], dtype=torch.int32, device=self.dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                settled = int(flag.item()) == 0
            if settled:
                break
        if self.world > 1:
            dist.barrier()
        launches0 = self.engine.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.perf_counter()
        for k in range(steps):
            flush.zero_()                                   # L2 flush between timed iterations (not timed)
            ev[k][0].record()
            self.device_step(n_mcmc)
            if k == steps - 1 and self.world > 1:
                self.flush_counts()                                  # nothing is left outside the timed region
                torch.cuda.current_stream().wait_stream(self.stats_stream)
            ev[k][1].record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        launches = self.engine.launch_count() - launches0 + 2 * steps       # + the two moment kernels per step
        step_ms = [a.elapsed_time(b) for a, b in ev]
        tmax = torch.tensor([sum(step_ms)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return tmax.item(), step_ms, launches, (t0, t1), len(extra)

    def roofline(self, pk):
        """The dominant kernel class (the dense-layer GEMM), timed live with CUDA events inside the library."""
        engine, B = self.engine, self.B
        prof = engine.profile_gemms(lambda: engine.local_energy(self.r))
        if not prof:
            return None
        fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        tensor_peak = pk["bf16_tflops"] / 6.0          # 3xTF32 = (bf16 / 2) / 3, SURVEY.md 8d
        ach = prof["flops"] / (prof["ms"] * 1e-3) / 1e12
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"       # dram__bytes_read.sum + dram__bytes_write.sum of the same launch shape (ncu --set full)
        if tf.exists():
            t = json.loads(tf.read_text()).get(f"{self.phys.name}:{B}:{prof['kernel']}")
            traffic = t["dram_bytes_per_launch"] if t else None
        flop_eval = FLOP_PER_EVAL.get(self.phys.name)
        return {"bound": "tensor", "kernel": prof["kernel"], "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": ach / tensor_peak, "traffic": traffic,
                "launch_shape": "largest-FLOP launches of the class (main embedding layers: rows = walkers x electrons x (3N+2), K = 320, N_out = 256)"
                                + ("; the CTA-pair kernel's launch also applies bias + spin-mean addend + tanh rule in its epilogue "
                                   "(the separate k_act pass it replaces is not counted as FLOPs)" if prof["kernel"] == "k_gemm_tc2_3xtf32" else ""),
                "class_achieved": prof["class_flops"] / (prof["class_ms"] * 1e-3) / 1e12, "class_launches": prof["class_count"],
                "class_frac": prof["class_flops"] / (prof["class_ms"] * 1e-3) / 1e12 / tensor_peak,
                "peak_source": f"{pk['source']}: bf16 {pk['bf16_tflops']} TF/s / 6 (3xTF32 FP32-accurate tensor peak); `frac` is against this BURST figure although the "
                               f"kernel is timed inside a 25 ms pass",
                "peak_sustained": pk["bf16_tflops_sustained"] / 6.0, "frac_sustained": ach / (pk["bf16_tflops_sustained"] / 6.0),
                "launches": prof["count"], "avg_launch_ms": prof["ms"] / max(prof["count"], 1),
                "algorithmic_flops_per_launch": prof["flops"] / max(prof["count"], 1),
                "frac_of_fp32_simt_peak": ach / fp32_peak, "fp32_simt_peak": fp32_peak,
                "share_of_eloc_time": prof["class_ms"] / prof["total_ms"],
                "eloc_pass_ms": prof["total_ms"], "eloc_stages_ms": prof.get("stages_ms"),
                "forward_stages_ms": {k: round(v[0], 4) for k, v in engine.profile_stages(lambda: engine.log_psi_sqr(self.r)).items()},
                "eloc_pass_frac": None if flop_eval is None else B * flop_eval / (prof["total_ms"] * 1e-3) / 1e12 / tensor_peak}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import deeperwin_b200 as dpe

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # the clock sampler is started before the warm-up: NVML start-up stalls the GPU for tens of ms and must not land in the timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    flush = torch.empty(512 * 2 ** 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    pk = peaks()
    B = args.walkers
    w = Workload(args, args.molecule, B, args.burn_in, world, rank, dev)
    n_el = w.phys.n_electrons
    dev_ms, step_ms, launches, (t_wall0, t_wall1), n_settle = w.timed(args.steps, args.warmup, flush)
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    value = B * world * args.steps / (dev_ms * 1e-3)

    # ---- the reference's cadence: n_inter_steps = 20 Metropolis steps per E_loc evaluation --------------------------------
    cadence = None
    if not args.no_cadence:
        n_inter = 20
        c_ms, c_steps, c_launches, _, _ = w.timed(3, 1, flush, n_mcmc=n_inter, settle=False)
        cadence = {"n_inter_steps": n_inter, "value": B * world * 3 / (c_ms * 1e-3), "unit": UNIT, "ms_per_epoch": c_ms / 3,
                   "metropolis_ms_per_step": (c_ms / 3 - dev_ms / args.steps) / (n_inter - 1), "gpu_launches": c_launches,
                   "what": "20 Metropolis steps + 1 forward-Laplacian E_loc + statistics per epoch (configuration.py:1039); evals/s = walkers / epoch time"}

        if w.engine.lib.dpe_kfac_layer_count(w.engine.handle) > 0:
            o_ms, o_launches = w.timed_epochs(3, n_inter, flush)
            cadence["optimisation_epoch"] = {
                "ms_per_epoch": o_ms, "value": B * world / (o_ms * 1e-3), "unit": UNIT, "gradient_and_kfac_ms": o_ms - c_ms / 3, "gpu_launches": o_launches,
                "what": "the epoch above + the parameter gradient of the clipped-energy loss and the KFAC factors (A, G) of every dense layer "
                        "(one backward pass, dpe_param_gradient; one flat all-reduce when N > 1); the optimiser's own update is outside the hot path"}

    # ---- end-to-end through the public Python API with host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        state, mc, log_psi_sqr, params, fixed = w.state, w.mc, w.log_psi_sqr, w.params, w.fixed
        n_up, n_dn = w.spin
        host = {k: getattr(state, k)[0].cpu().pin_memory() for k in ("r", "log_psi_sqr", "walker_age", "rng_state")}
        out_host = {k: torch.empty_like(v).pin_memory() for k, v in host.items()}
        e_host = torch.empty(B, dtype=torch.float32).pin_memory()
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = h2d + e_host.numel() * 4

        def api_step():
            s_dev = dpe.MCMCState(r=host["r"].to(dev, non_blocking=True)[None], R=state.R, Z=state.Z,
                                  log_psi_sqr=host["log_psi_sqr"].to(dev, non_blocking=True)[None],
                                  walker_age=host["walker_age"].to(dev, non_blocking=True)[None],
                                  rng_state=host["rng_state"].to(dev, non_blocking=True)[None],
                                  stepsize=state.stepsize, step_nr=state.step_nr, acc_rate=state.acc_rate, _step_nr_host=0)
            s_new = mc.run_inter_steps(log_psi_sqr, s_dev, params, n_up, n_dn, fixed)
            e = w.get_local_energy(params, (n_up, n_dn), s_new.r[0], w.R_, w.Z_, fixed)
            for k in out_host:
                out_host[k].copy_(getattr(s_new, k)[0], non_blocking=True)
            e_host.copy_(e, non_blocking=True)
            torch.cuda.synchronize()
            return float(e_host[0])

        for _ in range(max(1, args.warmup)):
            api_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            api_step()
        if world > 1:
            dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * args.steps / t.item(), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "MetropolisHastingsMonteCarlo.run_inter_steps(n_inter_steps=1) + get_local_energy, pinned host state in/out"}

    roofline = w.roofline(pk) if rank == 0 else None
    E_mean, acc_rate = float(w.aux["E_mean"]), float(w.ar.item())
    gemm_path = w.engine.lib.dpe_get_gemm_path(w.engine.handle)

    # ---- secondary block: the other molecule of BASELINE.json's metric (Benzene x 4096 walkers per GPU, configs[3]) ---------
    secondary = None
    if args.secondary and args.secondary != args.molecule:
        del w
        torch.cuda.empty_cache()
        w2 = Workload(args, args.secondary, B, min(args.burn_in, 20), world, rank, dev)
        s_ms, s_steps, s_launches, (s0, s1), _ = w2.timed(args.secondary_steps, 1, flush, settle=False)
        flop2 = FLOP_PER_EVAL.get(args.secondary)
        v2 = B * world * args.secondary_steps / (s_ms * 1e-3)
        secondary = {"metric": METRIC, "value": v2, "unit": UNIT, "n_gpus": world, "steps": args.secondary_steps, "warmup": 1,
                     "ms_per_step": s_ms / args.secondary_steps, "step_ms": [round(t, 2) for t in s_steps],
                     "config": {"workload": workload_string(args.secondary, w2.phys.n_electrons, B).replace("100 burn-in", f"{min(args.burn_in, 20)} burn-in"),
                                "walkers_per_gpu": B, "l2": "512 MiB flush between timed steps",
                                "chunks": "E_loc pass in workspace-sized chunks (computation.workspace_gb = 48)"},
                     "gpu_launches": s_launches, "clocks": sampler.window(s0, s1) if sampler else None,
                     "roofline": w2.roofline(pk) if rank == 0 else None,
                     "algorithmic_tflops": None if flop2 is None else v2 * flop2 / 1e12,
                     "E_mean": float(w2.aux["E_mean"]), "acc_rate": float(w2.ar.item())}
        del w2
    # ---- weight sharing (BASELINE.json configs[2]): one set of weights, 16 H10 geometries, ONE geometry per optimisation step ------
    shared = None
    if not args.no_weight_sharing:
        shared = weight_sharing_block(torch, world, dev, flush)
    if sampler:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flop_eval = FLOP_PER_EVAL.get(args.molecule)
    cpu = None if args.no_cpu_baseline else cpu_port_throughput(args.molecule, args.cpu_seconds)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args.molecule, n_el, B),
                       "walkers_per_gpu": B, "l2": "512 MiB flush between timed steps", "gemm_path": gemm_path,
                       "wall_ms_per_step_incl_flush": 1e3 * (t_wall1 - t_wall0) / args.steps},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "algorithmic_tflops": None if flop_eval is None else value * flop_eval / 1e12,
            "step_frac_of_tensor_peak": None if flop_eval is None else value / world * flop_eval / 1e12 / (pk["bf16_tflops"] / 6.0),
            "cadence": cadence, "secondary": secondary, "weight_sharing": shared,
            "E_mean": E_mean, "acc_rate": acc_rate, "step_ms": [round(t, 2) for t in step_ms],
            "extra_warmup_steps": n_settle}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def weight_sharing_block(torch, world, dev, flush, n_geom=16, walkers=512, n_inter=20):
    """configs[2]: H10 chains at 16 bond lengths sharing one set of weights (variational_optimization.py:356-387 optimises ONE geometry per
    step, all GPUs sharing its walkers).  A step = switch geometry (device-side, no host round trip) + n_inter Metropolis steps + E_loc +
    clipped statistics + parameter gradient and KFAC factors (+ their flat all-reduce), through the public callables; one round = 16 steps."""
    import numpy as np
    import deeperwin_b200 as dpe
    n_at = 10
    mapping = list(range(0, n_at, 2)) + list(range(1, n_at, 2))
    phys = [dpe.PhysicalConfig(name=f"HChain{n_at}_{a:.3f}", R=[[float(a) * k, 0.0, 0.0] for k in range(n_at)], Z=[1] * n_at, n_electrons=n_at,
                               n_up=n_at // 2, el_ion_mapping=mapping) for a in np.linspace(1.2, 3.6, n_geom)]
    cfg = dpe.Configuration(physical=phys[0].model_dump())
    f, _, _, params, fixed = dpe.build_log_psi_squared(cfg.model, phys[0], None, None, rng_seed=11, device=dev)
    gle = dpe.build_local_energy(f, forward_lap=True)
    vag = dpe.build_value_and_grad_func(f, gle, dpe.ClippingConfig(), with_kfac_statistics=True)
    mc = dpe.MetropolisHastingsMonteCarlo(dpe.MCMCConfigOptimization(n_inter_steps=n_inter, initialization="gaussian"))
    rank = int(os.environ.get("RANK", 0))
    spin = (n_at // 2, n_at // 2)
    geoms = [dpe.GeometryDataStore(idx=g, physical_config=p, spin_state=spin, fixed_params=fixed, clipping_state=dpe.init_clipping_state(),
                                   mcmc_state=dpe.MCMCState.initialize_around_nuclei(walkers, p, "gaussian", "el_ion_mapping", dpe.PRNGKey(100 * rank + g), device=dev))
             for g, p in enumerate(phys)]
    ema = {m: {k: v.clone() for k, v in l.items()} for m, l in params.items()}
    flat_p = [t for l in params.values() for t in l.values()]
    state = {"params": params, "epoch": 0}

    def sgd(p, grads, aux, lr=1e-4):          # stands in for the optimiser's update (KFAC preconditioning is outside the hot path): one multi-tensor kernel
        torch._foreach_add_(flat_p, [grads[m][k] for m, l in p.items() for k in l], alpha=-lr)
        return p

    def one_round():
        e = []
        for _ in range(n_geom):
            state["params"], idx, loss = dpe.shared_optimization_step(state["epoch"], geoms, f, vag, mc, state["params"], sgd, scheduling_method="round_robin",
                                                                      permutation=list(range(n_geom)), ema_params=ema)
            state["epoch"] += 1
            e.append(loss)
        return e

    for _ in range(3):                       # warm-up: burn-in of every geometry; graphs of the repeating Metropolis calls are captured
        one_round()
    torch.cuda.synchronize()
    launches0 = f.engine.launch_count()
    flush.zero_()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    e = one_round()
    a1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    return {"value": walkers * world * n_geom / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / n_geom, "steps": n_geom, "n_geometries": n_geom,
            "walkers_per_gpu_per_geometry": walkers, "n_inter_steps": n_inter, "gpu_launches": f.engine.launch_count() - launches0,
            "E_mean_first_last": [float(e[0]), float(e[-1])],
            "what": "H10 chain, 16 bond lengths, one shared set of weights, round-robin schedule (shared_optimization_step); step = geometry switch + "
                    "20 Metropolis steps + forward-Laplacian E_loc + statistics + gradient and KFAC factors (one flat all-reduce when N > 1) + an SGD "
                    "parameter update + EMA; evals/s = walkers x geometries / time of one round"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
