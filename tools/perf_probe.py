"""Throughput of the path kernel on the BASELINE shapes (development probe, not the bench contract).
usage: perf_probe.py [scale]   HEXO_GPU_LIB=<variant .so> selects the library under test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib

lib = _lib.load()
_lib.check(lib.hexo_gpu_init(0))
p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
stiff = hx.HParams(0.04, 0.04, -0.95, 20.0, 1.0)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive); E = hx.HQEAnderson(hx.EuropeanCallNonAdaptive)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
def run(name, scheme, T, K, n, steps, mode, p=p, rng="shishua", reps=3, drift="reference"):
    ch = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
    best = None
    for i in range(reps):
        r = hx.price_full(scheme, p, 100.0, ch, n, None, steps, seed=1, normal_mode=mode, rng=rng,
                          drift=drift)
        if i and (best is None or r.kernel_ms < best.kernel_ms): best = r
    r = best
    print(f"{name:24s} {mode} {rng:7s} n={n:.0e} steps={steps:4d} ms={r.kernel_ms:8.2f} "
          f"rate={r.path_steps/r.kernel_ms/1e6:7.2f} G/s price0={r.prices[0]:.4f}+-{r.stderr[0]:.4f}", flush=True)
print("lib:", _lib.LIB_PATH)
run("cfg4 asian 1024", A, [1.0], [[100.0]], int(2e7*scale), 1024, "f32")
if len(sys.argv) > 2 and sys.argv[2] == "short":   # A/B comparisons: the benchmark shape + two more
    run("cfg2 euro 252", E, [1.0], [[100.0]], int(4e6*scale), 252, "f32")
    run("cfg1-shape asian 252", A, [1.0], [[100.0]], int(4e6*scale), 252, "f32")
    run("cfg5 stiff 2520", A, [10.0], [list(np.linspace(70,130,64))], int(2e6*scale), 2520, "f32", stiff)
    run("cfg4 asian 1024", A, [1.0], [[100.0]], int(2e7*scale), 1024, "f64")
    run("cfg5 stiff 2520", A, [10.0], [list(np.linspace(70,130,64))], int(2e6*scale), 2520, "f64", stiff)
    run("cfg2 euro 252", E, [1.0], [[100.0]], int(4e6*scale), 252, "f64")
    sys.exit(0)
run("cfg2 euro 252", E, [1.0], [[100.0]], int(4e6*scale), 252, "f32")
run("cfg1-shape asian 252", A, [1.0], [[100.0]], int(4e6*scale), 252, "f32")
run("cfg5 stiff 2520", A, [10.0], [list(np.linspace(70,130,64))], int(2e6*scale), 2520, "f32", stiff)
run("cfg3 chain 64x8", A, [0.25*k for k in range(1,9)], [list(np.linspace(70,130,64))]*8, int(2e6*scale), 252, "f32")
run("cfg4 asian 1024", A, [1.0], [[100.0]], int(2e7*scale), 1024, "f64")
run("cfg4 asian 1024", A, [1.0], [[100.0]], int(2e7*scale), 1024, "f32", rng="philox")
run("cfg4 asian martingale", A, [1.0], [[100.0]], int(2e7*scale), 1024, "f32", drift="martingale")
run("cfg1 asian 100k", A, [1.0], [[100.0]], 100_000, 252, "f32", reps=4)

# cfg1 (the reference's own CLI size): how the stream count fills the machine
for ns in (0, 65536, 50000, 37888):
    ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
    best = None
    for i in range(4):
        r = hx.price_full(A, p, 100.0, ch, 100_000, 1, 252, seed=1, n_streams=ns)
        if i and (best is None or r.kernel_ms < best.kernel_ms): best = r
    print(f"cfg1 100k paths n_streams={best.n_streams:6d} grid={best.grid} ms={best.kernel_ms:.3f} "
          f"rate={best.path_steps/best.kernel_ms/1e6:7.2f} G/s", flush=True)
# cfg3 as SURVEY 8(d) defines it: 8 independent single-maturity pricings (64 strikes, 1e7 paths x 252
# steps each), submitted as one batch through the C ABI
import ctypes as C
from hestonexotics_b200 import pricing
K64 = list(np.linspace(70, 130, 64))
n3 = int(1e7 * min(scale, 1.0))
rqs = [pricing._Request(A, p, 100.0, [hx.OptionsChain.from_strikes(0.25 * k, K64)], n3, 64, 252, 1, "f32", 0)
       for k in range(1, 9)]
arr = (_lib.HexoPriceRequest * 8)(*[r.req for r in rqs])
prices, se = np.zeros(8 * 64), np.zeros(8 * 64)
stats = (_lib.HexoGpuStats * 8)()
for lanes in (1, 8):
    for rep in range(2):
        _lib.check(lib.hexo_gpu_price_batch(arr, 8, lanes, prices.ctypes.data_as(_lib.c_double_p),
                                            se.ctypes.data_as(_lib.c_double_p), stats))
    ms = stats[0].kernel_ms
    print(f"cfg3 as 8 independent pricings (batch, {lanes} lane(s)): n={n3:.0e} x 252 x 8 ms={ms:8.2f} "
          f"rate={8 * n3 * 252 / ms / 1e6:7.2f} G/s  ATM prices {prices.reshape(8, 64)[:, 32].round(4)}", flush=True)
