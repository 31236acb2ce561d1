#!/usr/bin/env python
"""Headline benchmark: gradient evals/sec (and min-ESS/sec) of the many-chain WALNUTS hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 1000-d ill-conditioned diagonal Gaussian,
sigma = logspace(-2, 2, 1000), 65 536 chains PER GPU (weak scaling), WALNUTSpy driver with
adaptLeapFrogR2P, H0 = 0.5, delta = 0.3, M = 10, minC = 0, maxC = 10.
One "step" = `--iters` transitions of every chain (one persistent-kernel launch).

Prints ONE JSON line (rank 0).  `value` = gradient evaluations per second with the chain states
resident in HBM, timed with CUDA events on the handle's stream (max over ranks); `e2e` = the same
through the public host-buffer API (H2D of the positions from pinned memory, D2H of draws,
diagnostics and positions inside the timed region).  `--impl reference` times the CPU restatement of
the reference Python implementation (oracle/walnutspy_oracle.py; the reference itself is Python and
cannot travel to the GPU box) on all host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = 1000
CHAINS_PER_GPU = 65536
CFG = dict(integrator="R2P", H0=0.5, delta=0.3, M=10, minC=0, maxC=10, jitter=0.2)
# FP64 work per gradient evaluation (= one leapfrog micro-step) and coordinate.  The reference's formulation is
# v+=a*g, q+=h*v, g=-q*s, v+=a*g = 4 FP64 instructions = 8 flop, 12 with the per-step energy SURVEY.md 8(d) quotes.
# The algorithm itself needs less: the closing and the opening half kick of neighbouring steps use the same gradient
# and merge, and the gradient of this target is linear, so an interior step is q+=h*v, v+=(-h*s)*q = 2 FMAs = 4 flop
# (energies are only consumed at the end of a pass).  The kernel executes exactly that, so 4 flop per coordinate is
# both the algorithmic minimum and the executed work; it is the roofline numerator.  The figures at the reference's
# 8 / 12 flop per coordinate are reported next to it.
FLOP_PER_DIM_PER_EVAL = 4
SEED = 20251017
MONITOR = 16                      # coordinates monitored for ESS (spread over the sigma range)


def sigma_vec():
    # monitored coordinates first: the sigma ordering is arbitrary for a diagonal target, so the
    # logspace is permuted such that the leading MONITOR coordinates span the whole range
    s = np.logspace(-2, 2, D)
    idx = np.unique(np.round(np.linspace(0, D - 1, MONITOR)).astype(int))
    rest = np.setdiff1d(np.arange(D), idx)
    return np.concatenate([s[idx], s[rest]])


def init_positions(n, rank, sigma):
    """q0 = sigma * z: exact draws from the target, so ESS is measured at stationarity."""
    rng = np.random.Generator(np.random.Philox(key=SEED + 7919 * rank))
    return rng.standard_normal((n, D)) * sigma


# --------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baseline: the numpy restatement of the reference, one chain per process on every host core
# --------------------------------------------------------------------------------------------------
def _cpu_chain(args):
    chain, iters = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import targets as ot
    from oracle import walnutspy_oracle as wo
    sigma = sigma_vec()
    lp = ot.make_diag_gauss(sigma)
    q0 = init_positions(chain + 1, 0, sigma)[chain]
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        s, dg = wo.WALNUTS(lp, q0, integrator=wo.ADAPT_R2P, H0=CFG["H0"], delta0=CFG["delta"], numIter=iters,
                           M=CFG["M"], igrAux=wo.AuxPar(CFG["minC"], CFG["maxC"]), seed=SEED, chain=chain)
    dt = time.perf_counter() - t0
    return float(dg[:, 6].sum() + dg[:, 7].sum()), dt


def cpu_baseline(iters=1, cores=None):
    """Gradient evals/sec of the reference algorithm (numpy port) on all host cores."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_chain, [(c, iters) for c in range(cores)])
    wall = time.perf_counter() - t0
    evals = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return {"value": evals / busy, "unit": "grad_evals/s", "cores": cores, "kind": "port",
            "sample": f"{cores} chains x {iters} transition(s) of the bench workload, one chain per process "
                      f"(numpy restatement of WALNUTSpy; {evals:.0f} evals in {busy:.1f}s, wall {wall:.1f}s)",
            "per_core": evals / busy / cores, "seconds": busy}


def cpu_baseline_c(iters=1, cores=None, chains_per_core=4):
    """The same workload on the C restatement (oracle/c/walnuts_oracle.c, -O3, pthreads): the strong CPU
    baseline standing in for the absent walnuts_cpp."""
    from oracle import c_oracle
    cores = cores or os.cpu_count() or 1
    sigma = sigma_vec()
    n = cores * chains_per_core
    q = np.ascontiguousarray(init_positions(n, 0, sigma))
    t0 = time.perf_counter()
    evals = c_oracle.run_many("diag_gauss", "R2P", q, CFG["H0"], CFG["delta"], CFG["M"], iters, SEED, cores,
                              minC=CFG["minC"], maxC=CFG["maxC"], inv_var=1.0 / sigma ** 2, jitter=CFG["jitter"])
    dt = time.perf_counter() - t0
    return {"value": evals / dt, "unit": "grad_evals/s", "cores": cores, "kind": "port-c",
            "sample": f"{n} chains x {iters} transition(s) on {cores} pthreads (C restatement, gcc -O3 -mavx2; "
                      f"{evals} evals in {dt:.1f}s)", "per_core": evals / dt / cores}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=2, help="transitions per chain per step")
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-iters", type=int, default=1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "diag_gauss_d1000_sigma_logspace(-2,2)_R2P", "chains_per_gpu": args.chains,
              "d": D, "iters_per_step": args.iters, **CFG,
              "l2_policy": "per-step working set (positions 524 MB + scratch) exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = None
        vals, secs = [], []
        for _ in range(max(1, min(args.steps, 3))):
            cb = cpu_baseline(args.cpu_iters)
            vals.append(cb["value"])
            secs.append(cb["seconds"])
        cb["value"] = float(np.mean(vals))
        try:
            extra_c = cpu_baseline_c(args.cpu_iters)
        except Exception as e:
            extra_c = {"unavailable": str(e)[:200]}
        line = {"impl": "reference", "cpu_baseline_c": extra_c, "metric": "grad_evals_per_sec", "value": cb["value"], "unit": "grad_evals/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * float(np.mean(secs)),     # one step = the bounded sample described in cpu_baseline
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "grad_evals/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from walnuts_b200 import ChainBatch, diagnostics, fp64_peak, _ffi

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    sigma = sigma_vec()
    n = args.chains
    q0 = init_positions(n, rank, sigma)
    cb = ChainBatch("diag_gauss", D, n, integrator=CFG["integrator"], H0=CFG["H0"], jitter=CFG["jitter"],
                    delta=CFG["delta"], M=CFG["M"], minC=CFG["minC"], maxC=CFG["maxC"], seed=SEED,
                    chain_offset=rank * n, device=local_rank, dg=MONITOR, data={"inv_var": 1.0 / sigma ** 2})
    cb.set_state(q0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    total_iters = args.iters * args.steps
    draws = torch.empty((total_iters, n, MONITOR), dtype=torch.float64, device=dev)
    for _ in range(args.warmup):
        cb.run_device(args.iters, draws=draws[:args.iters])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    kernel_ms, evals = [], 0
    for s in range(args.steps):
        cb.run_device(args.iters, draws=draws[s * args.iters:(s + 1) * args.iters], sync=False)
        cb.sync()
        kernel_ms.append(cb.last_kernel_ms())
        f, b = cb.last_grad_evals()
        evals += f + b
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(sum(kernel_ms))

    # ---- e2e: host buffers through the public API ------------------------------------------------
    host_q = torch.empty((n, D), dtype=torch.float64).pin_memory()
    cb.get_state(host_q.numpy())
    e2e_evals = 0
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = 0
    for s in range(args.steps):
        cb.set_state(host_q.numpy())
        out = cb.run(args.iters, draws=True, diag=True)
        cb.get_state(host_q.numpy())
        f, b = cb.last_grad_evals()
        e2e_evals += f + b
        h2d = host_q.numel() * 8
        d2h = out["draws"].nbytes + out["diag"].nbytes + out["nevalF"].nbytes + out["nevalB"].nbytes + host_q.numel() * 8
    barrier()
    e2e_wall = time.perf_counter() - t1

    # ---- reduce over ranks -----------------------------------------------------------------------
    stats = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(evals), float(e2e_evals)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_max, e2e_wall_max = [float(x) for x in stats]
    evals_all, e2e_evals_all = [float(x) for x in sums]

    # ---- min-ESS over the monitored coordinates (cross-chain, one small reduction over NVLink) ------
    ess_vals = []
    for j in range(MONITOR):
        z = draws[:, :, j].t().contiguous()
        st = diagnostics.chain_stats(z, max_lag=min(total_iters - 1, 32))
        vec = torch.stack([torch.as_tensor(float(st["m"]), device=dev, dtype=torch.float64), st["sum_mean"],
                           st["sum_mean2"], st["sum_var"], *st["acov_sum"]])
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM)
        vec = vec.cpu().numpy()
        st2 = dict(m=vec[0], n=total_iters, sum_mean=vec[1], sum_mean2=vec[2], sum_var=vec[3], acov_sum=vec[4:])
        ess_vals.append(diagnostics.ess_from_stats(st2)[0])
    min_ess = float(np.nanmin(ess_vals)) if total_iters >= 4 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = evals_all / (dev_ms_max * 1e-3)
    flops = FLOP_PER_DIM_PER_EVAL * D * (evals / max(1, args.steps))       # per launch (this rank)
    ach = flops / (np.mean(kernel_ms) * 1e-3) / 1e12
    try:
        peak = fp64_peak(local_rank) / 1e12
        peak_src = "measured FP64 FMA micro-benchmark (wn_fp64_peak) on this GPU, this run"
    except Exception:
        peak, peak_src = 148 * 64 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 64 FMA/clk x 1.965 GHz (fallback)"
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    line = {
        "metric": "grad_evals_per_sec", "value": value, "unit": "grad_evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "min_ess_per_sec": (min_ess / (dev_ms_max * 1e-3)) if min_ess else None,
        "min_ess": min_ess, "grad_evals": evals_all, "wall_s": wall_max,
        "e2e": {"value": e2e_evals_all / e2e_wall_max, "unit": "grad_evals/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": args.steps * cb.last_launches(),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     # dram__bytes_read+write of one `ncu --set full` launch (0.170 + 0.768 GB for 3.206e8 evaluations,
                     # profiles/r01_walnutspy_diag1000_R2P.txt), scaled to the evaluations of one bench launch
                     "traffic": 0.938e9 * (evals / max(1, args.steps)) / 3.206e8,
                     "flop_per_eval": FLOP_PER_DIM_PER_EVAL * D,
                     "achieved_at_8_flop_per_coord": ach * 8 / FLOP_PER_DIM_PER_EVAL,
                     "achieved_at_12_flop_per_coord": ach * 12 / FLOP_PER_DIM_PER_EVAL,
                     "peak_source": peak_src,
                     # SURVEY.md 8(d) streaming model: 48 d bytes per evaluation if (q, v, g) were re-read and
                     # re-written every micro-step; what HBM would have to deliver at the measured rate
                     "hbm": {"streaming_model_gbs": value / max(1, world) * 48 * D / 1e9, "peak_gbs": hbm_peak,
                             "streaming_model_over_peak": value / max(1, world) * 48 * D / 1e9 / hbm_peak},
                     "note": "register-resident chains: FP64 FMA pipe bound, not HBM (SURVEY.md 8d); "
                             f"HBM peak {hbm_peak} GB/s is not the limiter"},
        "lib": os.path.relpath(_ffi.lib_path(), ROOT),
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args.cpu_iters)
        if min_ess:
            # identical transition kernel on identical streams (parity tests) => identical ESS per gradient
            # evaluation; the CPU's min-ESS/sec is therefore its evals/s times the measured ESS per evaluation
            per_eval = min_ess / evals_all
            line["min_ess_per_grad_eval"] = per_eval
            line["cpu_baseline"]["min_ess_per_sec"] = per_eval * line["cpu_baseline"]["value"]
        try:
            line["cpu_baseline_c"] = cpu_baseline_c(args.cpu_iters)
            if min_ess:
                line["cpu_baseline_c"]["min_ess_per_sec"] = min_ess / evals_all * line["cpu_baseline_c"]["value"]
        except Exception as e:                                   # the C checker is optional for the bench
            line["cpu_baseline_c"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
