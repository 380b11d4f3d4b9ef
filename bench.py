#!/usr/bin/env python
"""bench.py -- Heston path-steps/s of the B200-native Monte-Carlo hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): one
arithmetic Asian call, K=S=100, T=1, kappa=2 theta=0.04 sigma=0.5 rho=-0.7 v0=0.04,
1e9 paths x 1024 steps, sharded over the N GPUs by disjoint RNG streams with one
all-reduce of the payoff sums ("scaling": "strong" -- the job is fixed, the shard
shrinks with N).  A "step" of this bench is one complete pricing of that job.

value  = path-steps/s with the request resident in HBM (prepared plan; the timed region is
         K x [path kernel + reduction kernel (+ NCCL all-reduce)], CUDA events, max over ranks).
e2e    = the same job through the public API hx.price_full / hx.price_distributed: host
         request buffers in, host prices out, every step.
--impl reference times the reference's own CPU code (oracle/_ref, built from
/root/reference/src by oracle/Makefile) on all host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARAMS = (0.04, 0.04, -0.7, 2.0, 0.5)
S0, STRIKE, EXPIRY, STEPS = 100.0, 100.0, 1.0, 1024
FULL_PATHS = 1_000_000_000
FLOP_PER_PATH_STEP = 100.0  # SURVEY.md section 8(d), Asian


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=FULL_PATHS, help="total paths of the job")
    ap.add_argument("--normal-mode", default="f32", choices=["f32", "f64"],
                    help="f32 = inverse normal as the reference is built (as241.f90:20-25)")
    ap.add_argument("--cpu-sample-paths", type=int, default=0,
                    help="paths of the CPU baseline sample (0 = sized for ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw else sm
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(sample_paths: int, steps: int):
    """The reference's own price<> (oracle/_ref) on all host cores.  Returns
    (path-steps/s, cores, seconds, kind, price)."""
    import oracle_api as oa
    c = oa.Contract(oa.ASIAN, [EXPIRY], [[STRIKE]], steps, PARAMS, S0)
    cores = os.cpu_count() or 1
    if oa.have_ref():
        r = oa.ref()
        t0 = time.perf_counter()
        price = c.ref_price(sample_paths, threads=cores)
        dt = time.perf_counter() - t0
        return sample_paths * steps / dt, int(r.ref_max_threads()), dt, "reference", float(price[0])
    # fallback: the single-threaded plain-C port
    t0 = time.perf_counter()
    pr, _, _ = c.price_ref(sample_paths, 1)
    dt = time.perf_counter() - t0
    return sample_paths * steps / dt, 1, dt, "port", float(pr[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # ~1e8 path-steps/s on 16 cores -> 2e5 paths x 1024 steps ~ 2 s per step
    sample = args.cpu_sample_paths or 12_500 * cores
    for _ in range(max(args.warmup, 0) and 1):
        cpu_reference_rate(max(sample // 8, cores), STEPS)
    t0 = time.perf_counter()
    rates, price = [], None
    for _ in range(args.steps):
        rate, used, dt, kind, price = cpu_reference_rate(sample, STEPS)
        rates.append(rate)
    total = time.perf_counter() - t0
    value = sample * STEPS * args.steps / total
    print(json.dumps({
        "impl": "reference", "metric": "heston_path_steps_per_sec", "value": value,
        "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg4 sample: Asian call K=100 T=1, {sample} paths x {STEPS} steps per "
                               "step on the host CPU (full job: 1e9 paths)", "price": price},
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": used, "kind": kind,
                         "sample": f"{sample} paths x {STEPS} steps x {args.steps} repeats"},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the pricing path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version / debug lines to stdout by default: keep stdout to the one
        # JSON line of the contract.  NCCL_DEBUG_FILE is only honoured above the VERSION level.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.hexo_gpu_init(local_rank))

    scheme = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
    p = hx.HParams(*PARAMS)
    chains = [hx.OptionsChain.from_strikes(EXPIRY, [STRIKE])]
    n_paths, n_opts = int(args.paths), 1
    n_streams = int(lib.hexo_gpu_default_streams(n_paths, n_opts, world))
    rq = hx.pricing._Request(scheme, p, S0, chains, n_paths, n_opts, STEPS, 1, args.normal_mode,
                             n_streams)
    begin, count = hx.shard_range(n_streams, rank, world)

    # ---- FP64 pipe peak (the roofline denominator is not in MEASURED_PEAKS.json) -------------
    fl, pk_ms = C.c_double(), C.c_float()
    _lib.check(lib.hexo_gpu_measure_fp64_peak(C.byref(fl), C.byref(pk_ms)))

    # ---- value: prepared plan, inputs resident ------------------------------------------------
    plan = C.c_void_p()
    _lib.check(lib.hexo_gpu_plan_create(C.byref(rq.req), begin, count, C.byref(plan)))
    stats = _lib.HexoGpuStats()
    _lib.check(lib.hexo_gpu_plan_stats(plan, C.byref(stats)))
    sums_t = torch.zeros(2 * n_opts, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def one_step():
        flush.zero_()
        _lib.check(lib.hexo_gpu_plan_launch(plan, C.c_void_p(sums_t.data_ptr()),
                                            C.c_void_p(stream.cuda_stream)))
        if world > 1:
            dist.all_reduce(sums_t)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        flush.zero_()
        kev[i][0].record()
        _lib.check(lib.hexo_gpu_plan_launch(plan, C.c_void_p(sums_t.data_ptr()),
                                            C.c_void_p(stream.cuda_stream)))
        kev[i][1].record()
        if world > 1:
            dist.all_reduce(sums_t)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
    t = torch.tensor([ms_total, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(t[0]), float(t[1])
    sums = sums_t.cpu().numpy()
    price = float(sums[0] / n_paths)
    se = float(np.sqrt(max(sums[1] / n_paths - price * price, 0.0) / n_paths))
    _lib.check(lib.hexo_gpu_plan_destroy(plan))

    path_steps = float(n_paths) * STEPS
    value = path_steps * args.steps / (ms_total * 1e-3)

    # ---- e2e: public API, host buffers in / host prices out, every step ------------------------
    def e2e_call():
        if world > 1:
            return hx.price_distributed(scheme, p, S0, chains, n_paths, n_opts, STEPS, seed=1,
                                        normal_mode=args.normal_mode, n_streams=n_streams)
        return hx.price_full(scheme, p, S0, chains, n_paths, n_opts, STEPS, seed=1,
                             normal_mode=args.normal_mode, n_streams=n_streams)
    e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        res = e2e_call()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = path_steps * args.e2e_steps / float(dt[0])
    h2d = rq.expiries.nbytes + rq.offsets.nbytes + rq.strikes.nbytes + C.sizeof(rq.req) \
        + 144 * len(rq.expiries)            # request, flattened chains, segment constants
    d2h = 2 * n_opts * 8

    if rank == 0:
        achieved = FLOP_PER_PATH_STEP * (path_steps / world) / (kernel_ms * 1e-3)  # this GPU
        out = {
            "metric": "heston_path_steps_per_sec", "value": value, "unit": "path-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"cfg4: arithmetic Asian call K=100 S=100 T=1, {n_paths} paths x {STEPS} "
                            f"steps total, sharded over {world} GPU(s) by RNG stream",
                "heston": dict(zip(("v0", "theta", "rho", "kappa", "sigma"), PARAMS)),
                "scheme": "Andersen QE psi_c=1.5, reference-compatible last-step rule",
                "rng": "shishua, one stream per thread, seed {1, stream, 0, 0}",
                "normal_mode": args.normal_mode,
                "n_streams": n_streams, "grid": int(stats.grid), "block": int(stats.block),
                "steps_per_path": int(stats.steps_per_path),
                "l2": "256 MiB memset before every step (inside the timed region); the kernel reads "
                      "<1 KB of input",
                "price": price, "stderr": se,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "path-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
                    "api": "hx.price_distributed" if world > 1 else "hx.price_full",
                    "price": float(res.prices[0])},
            "gpu_launches": 2 * args.steps,
            "roofline": {
                "bound": "fp64", "achieved": achieved / 1e12, "peak": fl.value / 1e12,
                "unit": "TFLOP/s", "frac": achieved / fl.value,
                # dram__bytes_read.sum + dram__bytes_write.sum of the path kernel in the round-1
                # `ncu --set full` capture (profiles/r01_path_kernel_ncu_raw_final.csv, a
                # 4.1e9-path-step launch): 85 248 B read, 0 B written -- code and constants only
                # (61 KB ... 469 KB from capture to capture), it does not grow with the paths
                "traffic": 85248,
                "note": "achieved = 100 algorithmic FP64 flop per path-step (SURVEY 8d) x path-steps "
                        "of one GPU / mean path-kernel time (CUDA events); peak = DFMA peak measured "
                        "in this run (hexo_gpu_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 "
                        "entry); bound is the FP64 ALU pipe (SURVEY 8d), not HBM or tensor: DRAM "
                        "traffic (bytes per launch, from ncu) is ~0 by design",
                "kernel": "heston_qe_paths_kernel", "kernel_ms": kernel_ms,
            },
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_sample_paths or 100_000 * cores
            rate, used, secs, kind, cprice = cpu_reference_rate(sample, STEPS)
            out["cpu_baseline"] = {
                "value": rate, "unit": "path-steps/s", "cores": used, "kind": kind,
                "sample": f"{sample} paths x {STEPS} steps of the same contract, {secs:.1f} s, "
                          f"price {cprice:.4f}"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
