"""Quick throughput probe used during development (not the bench contract)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib

lib = _lib.load()
_lib.check(lib.hexo_gpu_init(0))
fl = C.c_double(); ms = C.c_float()
_lib.check(lib.hexo_gpu_measure_fp64_peak(C.byref(fl), C.byref(ms)))
print(f"fp64 DFMA peak: {fl.value/1e12:.2f} TFLOP/s ({ms.value:.2f} ms)")
p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive); E = hx.HQEAnderson(hx.EuropeanCallNonAdaptive)
def run(name, scheme, T, K, n, steps, mode, p=p, rng="shishua"):
    ch = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
    r = hx.price_full(scheme, p, 100.0, ch, n, None, steps, seed=1, normal_mode=mode, rng=rng)
    r = hx.price_full(scheme, p, 100.0, ch, n, None, steps, seed=2, normal_mode=mode, rng=rng)
    print(f"{name:28s} {mode} n={n:.1e} steps={steps} ms={r.kernel_ms:9.2f} rate={r.path_steps/r.kernel_ms/1e6:8.2f} Gps/s grid={r.grid}x{r.block} price0={r.prices[0]:.4f}+-{r.stderr[0]:.4f}", flush=True)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
for mode in ("f32", "f64"):
    run("cfg1 asian 100k", A, [1.0], [[100.0]], 100_000, 252, mode)
    run("cfg2 euro 1M", E, [1.0], [[100.0]], 1_000_000, 252, mode)
    run("cfg4 asian 1024", A, [1.0], [[100.0]], int(2e7*scale), 1024, mode)
    run("cfg3 chain 64x8", A, [0.25*k for k in range(1,9)], [list(np.linspace(70,130,64))]*8, int(2e6*scale), 252, mode)
    run("cfg5 stiff", A, [10.0], [list(np.linspace(70,130,64))], int(2e6*scale), 2520, mode, hx.HParams(0.04,0.04,-0.95,20.0,1.0))
run("cfg4 asian 1024 PHILOX", A, [1.0], [[100.0]], int(2e7*scale), 1024, "f32", rng="philox")
run("cfg2 euro 1M PHILOX", E, [1.0], [[100.0]], 1_000_000, 252, "f32", rng="philox")
# batched submission: cfg1-sized jobs (10^5 paths x 252 steps) for 64 parameter sets
import time
ps = [hx.HParams(0.04 + 0.0005 * i, 0.04, -0.7, 2.0, 0.5) for i in range(64)]
ch1 = [hx.OptionsChain.from_strikes(1.0, [100.0])]
for _ in range(2):
    t0 = time.perf_counter()
    for q in ps:
        hx.price_full(A, q, 100.0, ch1, 100_000, None, 252)
    t_seq = time.perf_counter() - t0
for lanes in (1, 2, 4, 8, 16):
    for _ in range(2):
        t0 = time.perf_counter()
        pr, se, ms = hx.price_batch(A, ps, 100.0, ch1, 100_000, None, 252, n_lanes=lanes)
        t_b = time.perf_counter() - t0
    print(f"batch 64 x cfg1: lanes={lanes:2d} device {ms:7.2f} ms  host {t_b*1e3:7.2f} ms  "
          f"({64*100_000*252/t_b/1e9:6.1f} Gps/s)   sequential price_full {t_seq*1e3:7.2f} ms "
          f"({64*100_000*252/t_seq/1e9:6.1f} Gps/s)", flush=True)
