import sys, os
sys.path.insert(0, "/root/repo")
import hestonexotics_b200 as hx
p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive); E = hx.HQEAnderson(hx.EuropeanCallNonAdaptive)
def run(name, scheme, n, steps, mode, rng="shishua"):
    ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
    for sd in (1, 2):
        r = hx.price_full(scheme, p, 100.0, ch, n, None, steps, seed=sd, normal_mode=mode, rng=rng)
    print(f"{name:22s} {mode} {rng:8s} ms={r.kernel_ms:9.2f} rate={r.path_steps/r.kernel_ms/1e6:8.2f} Gps/s price={r.prices[0]:.4f}+-{r.stderr[0]:.4f}", flush=True)
run("cfg4 asian 1024", A, 20_000_000, 1024, "f32")
run("cfg4 asian 1024", A, 20_000_000, 1024, "f64")
run("cfg2 euro 252", E, 4_000_000, 252, "f32")
run("cfg4 asian 1024", A, 20_000_000, 1024, "f32", "philox")
run("cfg1 asian 252", A, 100_000, 252, "f32")
