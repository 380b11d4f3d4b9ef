#!/usr/bin/env python
"""bench.py -- Heston path-steps/s of the B200-native Monte-Carlo hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python bench.py --capi-multi N [--workload cfg4|cfg5]   # one process, N GPUs, C ABI only

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
--capi-multi N prices the job through hexo_gpu_price_multi: ONE process driving N GPUs, the
form the reference's single-process CLI would use (no torch.distributed involved).
--workload cfg5 switches to BASELINE.json configs[4] (stiff regime, 64 strikes, 1e8 paths x 2520
steps); the default is configs[3], the configuration the metric is quoted on.
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
# Measured constants of the path kernel (profiles/, ncu captures of the same kernel build; they do
# not depend on the job size): FP64 instructions issued per path-step, of which FMA, and the
# kernel's DRAM traffic per launch.  bench.py reports them next to the algorithmic figure.
# per normal mode: (FP64 instructions, of which DFMA) -- a DFMA counts two flop
FP64_INSTR = {"f32": (40.13, 28.14), "f32-ppnd7": (40.13, 28.14), "f64": (93.67, 73.00)}
DRAM_BYTES_PER_LAUNCH = {"f32": 68096, "f32-ppnd7": 63488, "f64": 98560}
NCU_PROFILES = {"f32": "profiles/r02_path_kernel_ncu_keys.txt",
                "f32-ppnd7": "profiles/r02_path_kernel_ncu_keys_ppnd7.txt",
                "f64": "profiles/r02_path_kernel_ncu_keys_f64.txt"}
# sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active of the path kernel per normal mode,
# from the ncu captures in profiles/ (r02_path_kernel_ncu_keys{,_ppnd7,_f64}.txt)
NCU_PIPE_FP64_PCT = {"f32": 37.7, "f32-ppnd7": 39.9, "f64": 63.7}
WORKLOADS = {
    # name: (params, expiry, strikes, steps, paths, description)
    "cfg4": (PARAMS, EXPIRY, [STRIKE], STEPS, FULL_PATHS,
             "cfg4: arithmetic Asian call K=100 S=100 T=1"),
    "cfg5": ((0.04, 0.04, -0.95, 20.0, 1.0), 10.0, [70.0 + 60.0 * i / 63 for i in range(64)], 2520,
             100_000_000, "cfg5: stiff regime (kappa=20 sigma=1 rho=-0.95), Asian chain of 64 "
                          "strikes, T=10, daily steps"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=0, help="total paths of the job (0 = the workload's)")
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--capi-multi", type=int, default=0, metavar="N",
                    help="one process, N GPUs through hexo_gpu_price_multi (no torchrun)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the supplementary legs (other normal modes, reference z-score)")
    ap.add_argument("--normal-mode", default="f32", choices=["f32", "f64", "f32-ppnd7"],
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


def reference_z_score(price, se, n_paths, steps, ref_paths=100_000):
    """The GPU price against the reference's own code on ONE host thread (race-free, ~10 s):
    z = (gpu - reference) / combined standard error; the reference keeps no standard error, so
    its side uses the per-path deviation the GPU measured for the same payoff."""
    import numpy as np
    import oracle_api as oa
    if not oa.have_ref():
        return None
    c = oa.Contract(oa.ASIAN, [EXPIRY], [[STRIKE]], steps, PARAMS, S0)
    ref = float(c.ref_price(ref_paths, threads=1)[0])
    se_ref = se * np.sqrt(n_paths / ref_paths)
    return {"z": (price - ref) / float(np.hypot(se, se_ref)), "gpu_price": price, "gpu_stderr": se,
            "reference_price": ref, "reference_stderr": float(se_ref), "reference_paths": ref_paths,
            "reference_threads": 1}


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

    params, expiry, strikes, steps, full_paths, what = WORKLOADS[args.workload]
    scheme = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
    p = hx.HParams(*params)
    chains = [hx.OptionsChain.from_strikes(expiry, strikes)]
    n_paths, n_opts = int(args.paths or full_paths), len(strikes)
    atm = int(np.argmin(np.abs(np.asarray(strikes) - S0)))
    n_streams = int(lib.hexo_gpu_default_streams(n_paths, n_opts, world))
    begin, count = hx.shard_range(n_streams, rank, world)

    # ---- FP64 pipe peak (the roofline denominator is not in MEASURED_PEAKS.json) -------------
    fl, pk_ms = C.c_double(), C.c_float()
    _lib.check(lib.hexo_gpu_measure_fp64_peak(C.byref(fl), C.byref(pk_ms)))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_plan(normal_mode, n_steps, n_warm, sample_clocks):
        """K launches of a prepared plan (inputs resident in HBM): returns total ms, mean path-kernel
        ms (CUDA events on the launching stream, max over ranks), the sums and the clocks."""
        rq = hx.pricing._Request(scheme, p, S0, chains, n_paths, n_opts, steps, 1, normal_mode,
                                 n_streams)
        plan = C.c_void_p()
        _lib.check(lib.hexo_gpu_plan_create(C.byref(rq.req), begin, count, C.byref(plan)))
        stats = _lib.HexoGpuStats()
        _lib.check(lib.hexo_gpu_plan_stats(plan, C.byref(stats)))
        sums_t = torch.zeros(2 * n_opts, dtype=torch.float64, device=dev)
        for _ in range(n_warm):
            flush.zero_()
            _lib.check(lib.hexo_gpu_plan_launch(plan, C.c_void_p(sums_t.data_ptr()),
                                                C.c_void_p(stream.cuda_stream)))
            if world > 1:
                dist.all_reduce(sums_t)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(n_steps)]
        ev0.record()
        for i in range(n_steps):
            flush.zero_()
            kev[i][0].record()
            _lib.check(lib.hexo_gpu_plan_launch(plan, C.c_void_p(sums_t.data_ptr()),
                                                C.c_void_p(stream.cuda_stream)))
            kev[i][1].record()
            if world > 1:
                dist.all_reduce(sums_t)
        ev1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 and sample_clocks else None
        t = torch.tensor([ev0.elapsed_time(ev1),
                          statistics.mean(a.elapsed_time(b) for a, b in kev)],
                         dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sums = sums_t.cpu().numpy()
        _lib.check(lib.hexo_gpu_plan_destroy(plan))
        return float(t[0]), float(t[1]), sums, clocks, stats, rq

    # ---- value: prepared plan, inputs resident ------------------------------------------------
    ms_total, kernel_ms, sums, clocks, stats, rq = timed_plan(args.normal_mode, args.steps,
                                                              args.warmup, True)
    price = float(sums[atm] / n_paths)
    se = float(np.sqrt(max(sums[n_opts + atm] / n_paths - price * price, 0.0) / n_paths))
    path_steps = float(n_paths) * steps
    value = path_steps * args.steps / (ms_total * 1e-3)

    # ---- e2e: public API, host buffers in / host prices out, every step ------------------------
    def e2e_call():
        if world > 1:
            return hx.price_distributed(scheme, p, S0, chains, n_paths, n_opts, steps, seed=1,
                                        normal_mode=args.normal_mode, n_streams=n_streams)
        return hx.price_full(scheme, p, S0, chains, n_paths, n_opts, steps, seed=1,
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
        + 152 * len(rq.expiries)            # request, flattened chains, segment constants
    d2h = 2 * n_opts * 8

    # ---- supplementary: the other normal modes on the same job (one timed launch each) ---------
    other_modes = {}
    if not args.no_extras:
        for mode in ("f32", "f32-ppnd7", "f64"):
            if mode == args.normal_mode:
                continue
            m_total, m_kernel, m_sums, _, _, _ = timed_plan(mode, 1, 1, False)
            other_modes[mode] = {"value": path_steps / (m_total * 1e-3),
                                 "price": float(m_sums[atm] / n_paths),
                                 "roofline_frac": FLOP_PER_PATH_STEP * path_steps / world /
                                 (m_kernel * 1e-3) / fl.value,
                                 "fp64_issued_frac": sum(FP64_INSTR[mode]) * path_steps / world /
                                 (m_kernel * 1e-3) / fl.value,
                                 "ncu_pipe_fp64_pct": NCU_PIPE_FP64_PCT[mode]}

    if rank == 0:
        this_gpu = path_steps / world / (kernel_ms * 1e-3)       # path-steps/s of one GPU
        achieved = FLOP_PER_PATH_STEP * this_gpu
        fp64_instr, fp64_fma = FP64_INSTR[args.normal_mode]
        issued = (fp64_instr + fp64_fma) * this_gpu
        ncu_profile = NCU_PROFILES[args.normal_mode]
        out = {
            "metric": "heston_path_steps_per_sec", "value": value, "unit": "path-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{what}, {n_paths} paths x {steps} steps total, sharded over {world} "
                            f"GPU(s) by RNG stream",
                "heston": dict(zip(("v0", "theta", "rho", "kappa", "sigma"), params)),
                "scheme": "Andersen QE psi_c=1.5, reference-compatible last-step rule",
                "rng": "shishua, one stream per thread, seed {1, stream, 0, 0}",
                "normal_mode": args.normal_mode + (
                    " (AS241 PPND16 evaluated in single precision, as the reference is built)"
                    if args.normal_mode == "f32" else ""),
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
                    "price": float(res.prices[atm]),
                    "note": f"{args.e2e_steps} timed calls (a call takes seconds at this size); "
                            f"`value` is timed over {args.steps}"},
            "gpu_launches": 2 * args.steps,
            "roofline": {
                "bound": "fp64", "achieved": achieved / 1e12, "peak": fl.value / 1e12,
                "unit": "TFLOP/s", "frac": achieved / fl.value,
                # what the FP64 pipe really executes (FP64_INSTR, from the ncu capture of this mode's
                # kernel): as-built F32 normals 40.2 FP64 instructions per path-step (28.1 FMAs),
                # not the 100 flop of the reference algorithm -- 66 of those belong to PPND16, which
                # that mode evaluates on the FP32 pipe; F64 normals 93.7 (73.0 FMAs)
                "fp64_issued_frac": issued / fl.value,
                "fp64_issued_flop_per_path_step": fp64_instr + fp64_fma,
                "fp64_instr_per_path_step": fp64_instr,
                "ncu_pipe_fp64_pct": NCU_PIPE_FP64_PCT.get(args.normal_mode),
                "ncu": "sm__inst_executed_pipe_fp64 and the instruction counts above are from the "
                       f"ncu capture of this kernel build, {ncu_profile}",
                # dram__bytes_read.sum + dram__bytes_write.sum of the path kernel per launch, from
                # the same capture (code and constants only: it does not grow with the paths) --
                # a profile figure, not a measurement of this run
                "traffic": DRAM_BYTES_PER_LAUNCH[args.normal_mode], "traffic_source": ncu_profile,
                "note": "achieved = 100 algorithmic FP64 flop per path-step (SURVEY 8d) x path-steps "
                        "of one GPU / mean path-kernel time (CUDA events); peak = DFMA peak measured "
                        "in this run (hexo_gpu_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 "
                        "entry); bound is the FP64 ALU pipe (SURVEY 8d), not HBM or tensor: DRAM "
                        "traffic is ~0 by design",
                "kernel": "heston_qe_paths_kernel", "kernel_ms": kernel_ms,
            },
        }
        if other_modes:
            out["other_normal_modes"] = {
                "note": "the same job, one timed launch per mode; f32-ppnd7 = AS241's single-"
                        "precision routine PPND7 (optional), f64 = PPND16 in double (all 100 "
                        "algorithmic flop on the FP64 pipe); ncu_pipe_fp64_pct from profiles/",
                **other_modes}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_sample_paths or 100_000 * cores
            rate, used, secs, kind, cprice = cpu_reference_rate(sample, STEPS)
            out["cpu_baseline"] = {
                "value": rate, "unit": "path-steps/s", "cores": used, "kind": kind,
                "sample": f"{sample} paths x {STEPS} steps of the cfg4 contract, {secs:.1f} s, "
                          f"price {cprice:.4f}"}
            if not args.no_extras and args.workload == "cfg4" and world == 1:
                out["price_z_vs_reference"] = reference_z_score(price, se, n_paths, steps)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_capi_multi(args):
    """One process, N GPUs, C ABI only: hexo_gpu_price_multi on the full job (the path the
    reference's single-process CLI would take).  Timed on the host around the call (it returns
    host prices), device time = the slowest GPU's CUDA-event time from the call's stats."""
    import numpy as np
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import _lib
    lib = _lib.load()
    n = args.capi_multi
    if lib.hexo_gpu_device_count() < n:
        raise SystemExit(f"bench.py --capi-multi {n}: only {lib.hexo_gpu_device_count()} device(s)")
    params, expiry, strikes, steps, full_paths, what = WORKLOADS[args.workload]
    scheme = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
    p = hx.HParams(*params)
    chains = [hx.OptionsChain.from_strikes(expiry, strikes)]
    n_paths, n_opts = int(args.paths or full_paths), len(strikes)
    atm = int(np.argmin(np.abs(np.asarray(strikes) - S0)))
    _lib.check(lib.hexo_gpu_init(0))
    n_streams = int(lib.hexo_gpu_default_streams(n_paths, n_opts, n))
    rq = hx.pricing._Request(scheme, p, S0, chains, n_paths, n_opts, steps, 1, args.normal_mode,
                             n_streams)
    prices, se = np.zeros(n_opts), np.zeros(n_opts)
    stats = _lib.HexoGpuStats()

    def call():
        _lib.check(lib.hexo_gpu_price_multi(C.byref(rq.req), n, prices.ctypes.data_as(_lib.c_double_p),
                                            se.ctypes.data_as(_lib.c_double_p), C.byref(stats)))
    for _ in range(max(args.warmup, 1)):
        call()
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = []
    for _ in range(args.steps):
        call()
        dev_ms.append(float(stats.kernel_ms))
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    path_steps = float(n_paths) * steps
    # the sums do not depend on how the streams are spread over devices: the same request on ONE
    # device (same n_streams) must give the same prices
    one = hx.price_full(scheme, p, S0, chains, n_paths, n_opts, steps, seed=1,
                        normal_mode=args.normal_mode, n_streams=n_streams) if n > 1 else None
    print(json.dumps({
        "impl": "ours-capi-multi", "metric": "heston_path_steps_per_sec",
        "value": path_steps * args.steps / (statistics.mean(dev_ms) * 1e-3 * args.steps),
        "unit": "path-steps/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{what}, {n_paths} paths x {steps} steps, {n} GPU(s) driven by one "
                               "process through hexo_gpu_price_multi",
                   "normal_mode": args.normal_mode, "n_streams": n_streams,
                   "price": float(prices[atm]), "stderr": float(se[atm])},
        "clocks": clocks,
        "e2e": {"value": path_steps * args.steps / dt, "unit": "path-steps/s",
                "api": "hexo_gpu_price_multi (C ABI, host request in, host prices out)"},
        "device_ms_slowest_gpu": statistics.mean(dev_ms),
        "same_prices_as_one_device": None if one is None else bool(
            np.allclose(prices, one.prices, rtol=1e-12, atol=0) and
            np.allclose(se, one.stderr, rtol=1e-9)),
        "gpu_launches": 2 * n * args.steps,
    }))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.capi_multi:
        run_capi_multi(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
