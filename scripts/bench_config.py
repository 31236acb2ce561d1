#!/usr/bin/env python
"""bench.py's measurement for the OTHER BASELINE.json configs (bench.py itself stays on configs[1], the config the
metric is quoted on): gradient evals/s and min-ESS/s with chains resident in HBM, CUDA events on the handle's
stream, max over ranks, weak scaling under torchrun.  One JSON line (rank 0).

    python scripts/bench_config.py --config c3|c4|c5 [--steps K --warmup W --iters I --chains N]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        scripts/bench_config.py --config c5          # 1 048 576 Stock-Watson chains over 8 GPUs
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402

from oracle import targets as ot   # noqa: E402  (data generators only: synthetic X, y; the Stock-Watson series)

CONFIGS = {
    # funnel10, mainFunnel.py:24-32 settings, exact funnel draws as initial states
    "c3": dict(target="funnel", d=11, chains=262144, integrator="R2P", H0=0.3, delta=0.3, M=12, minC=0, maxC=10,
               iters=10, monitor=11, flop_per_eval=200.0, workload="funnel10_R2P_M12"),
    # logistic regression N = 100 000, P = 100 (synthetic, seed 0)
    "c4": dict(target="logreg", d=100, chains=16384, integrator="R2P", H0=0.05, delta=0.3, M=6, minC=0, maxC=10,
               iters=1, monitor=8, flop_per_eval=4.5e7, workload="logreg_N100000_P100_R2P"),
    # Stock-Watson, mainSW.py:41-80 settings (M = 14, H0 = 0.1, delta0 = 0.3, minC = 3), 131 072 chains per GPU
    "c5": dict(target="stock_watson", d=756, chains=131072, integrator="R2P", H0=0.1, delta=0.3, M=14, minC=3,
               maxC=10, iters=1, monitor=8, flop_per_eval=4.0e4, workload="stock_watson_T252_R2P_M14_minC3"),
}


def make(cfg, n, rank):
    rng = np.random.Generator(np.random.Philox(key=1234 + 7919 * rank))
    if cfg["target"] == "funnel":
        q0 = np.empty((n, 11))
        q0[:, 0] = 3.0 * rng.standard_normal(n)
        q0[:, 1:] = np.exp(0.5 * q0[:, :1]) * rng.standard_normal((n, 10))
        return q0, {}
    if cfg["target"] == "logreg":
        X, y, beta = ot.synth_logreg_data(100_000, 100, 0)
        return beta + 0.05 * rng.standard_normal((n, 100)), {"X": X, "y": y, "tau": np.array([1.0])}
    y = ot.load_sw_data()
    q0 = 0.05 * rng.standard_normal((n, 756))
    q0[:, 0] = 2.4
    return q0, {"y": y}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(CONFIGS))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("--chains", type=int, default=0)
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    iters = a.iters or cfg["iters"]
    n = a.chains or cfg["chains"]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    from bench import ClockSampler
    from walnuts_b200 import ChainBatch, diagnostics, fp64_peak
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    q0, data = make(cfg, n, rank)
    mon = cfg["monitor"]
    cb = ChainBatch(cfg["target"], cfg["d"], n, integrator=cfg["integrator"], H0=cfg["H0"], delta=cfg["delta"],
                    M=cfg["M"], minC=cfg["minC"], maxC=cfg["maxC"], seed=20251017, chain_offset=rank * n,
                    device=local_rank, dg=mon, data=data)
    cb.set_state(q0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total = iters * a.steps
    draws = torch.empty((total, n, mon), dtype=torch.float64, device=dev)
    for _ in range(a.warmup):
        cb.run_device(iters, draws=draws[:iters])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, evals = 0.0, 0
    for s in range(a.steps):
        cb.run_device(iters, draws=draws[s * iters:(s + 1) * iters])
        ms += cb.last_kernel_ms()
        f, b = cb.last_grad_evals()
        evals += f + b
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    st = torch.tensor([ms], dtype=torch.float64, device=dev)
    sm = torch.tensor([float(evals)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    ms_max, evals_all = float(st[0]), float(sm[0])

    ess = []
    if total >= 4:
        for j in range(mon):
            z = draws[:, :, j].t().contiguous()
            c = diagnostics.chain_stats(z, max_lag=min(total - 1, 32))
            vec = torch.stack([torch.as_tensor(float(c["m"]), device=dev, dtype=torch.float64), c["sum_mean"],
                               c["sum_mean2"], c["sum_var"], *c["acov_sum"]])
            if world > 1:
                dist.all_reduce(vec, op=dist.ReduceOp.SUM)
            v = vec.cpu().numpy()
            ess.append(diagnostics.ess_from_stats(dict(m=v[0], n=total, sum_mean=v[1], sum_mean2=v[2], sum_var=v[3],
                                                       acov_sum=v[4:]))[0])
    if rank == 0:
        value = evals_all / (ms_max * 1e-3)
        peak = fp64_peak(local_rank) / 1e12
        ach = value / world * cfg["flop_per_eval"] / 1e12
        min_ess = float(np.nanmin(ess)) if ess else None
        print(json.dumps({
            "metric": "grad_evals_per_sec", "value": value, "unit": "grad_evals/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "chains_per_gpu": n, "chains_total": n * world, "d": cfg["d"],
                       "iters_per_step": iters, **{k: cfg[k] for k in ("integrator", "H0", "delta", "M", "minC", "maxC")}},
            "evals_per_transition": evals_all / (total * n * world),
            "min_ess": min_ess, "min_ess_per_sec": (min_ess / (ms_max * 1e-3)) if min_ess else None,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "flop_per_eval": cfg["flop_per_eval"], "note": "algorithmic flop per evaluation, SURVEY.md 8(d)"},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
