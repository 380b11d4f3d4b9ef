"""Bias of the QE scheme (reference drift, and martingale-corrected) on Andersen's test cases as
the reference's sketched simulation_test lists them (src/UnitTest.cpp:565-596): European calls,
strikes 70/100/140, bias = Monte-Carlo price - closed-form Heston price (r = 0).
usage: andersen_probe.py [n_paths]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import hestonexotics_b200 as hx
from heston_cf import heston_call

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
E = hx.HQEAnderson(hx.EuropeanCallNonAdaptive)
K = [70.0, 100.0, 140.0]
cases = [("sketch-0 T=5", (0.04, 0.04, -0.9, 0.4, 1.0), 5.0), ("sketch-1 T=10", (0.04, 0.04, -0.5, 0.3, 1.0), 10.0),
         ("sketch-2 T=15", (0.09, 0.09, -0.3, 1.0, 1.0), 15.0),
         ("andersen-I T=10", (0.04, 0.04, -0.9, 0.5, 1.0), 10.0), ("andersen-II T=15", (0.04, 0.04, -0.5, 0.3, 0.9), 15.0),
         ("andersen-III T=5", (0.09, 0.09, -0.3, 1.0, 1.0), 5.0)]
for name, p, T in cases:
    cf = np.array([heston_call(100.0, k, T, *p, r=0.0) for k in K])
    print(f"== {name} params {p}: closed form {cf}")
    for inv in (1, 2, 4, 8, 16, 32):
        steps = int(round(T * inv))
        for drift in ("reference", "martingale"):
            r = hx.price_full(E, hx.HParams(*p), 100.0, [hx.OptionsChain.from_strikes(T, K)], n, 3, steps,
                              seed=1, normal_mode="f64", drift=drift, time_grid="exact")
            print(f"  delta=1/{inv:<2d} {drift:10s} bias " + " ".join(f"{b:+8.3f}({s:.3f})" for b, s in zip(r.prices - cf, r.stderr)))
