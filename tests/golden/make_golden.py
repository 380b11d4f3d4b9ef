"""Generates tests/golden/*.json from the REFERENCE'S OWN SOURCES compiled into
oracle/_ref/libhexo_ref.so (oracle/Makefile; needs /root/reference, i.e. the
build container).  Run:  python tests/golden/make_golden.py

Each fixture is an output of the unmodified reference code
(HSimulation::price<...>, RNG::get_urand/get_grand) run single-threaded
(the multi-threaded accumulation is racy, src/HSimulation.tpp:40), with the
oracle's shishua / ppnd16 shims underneath (see oracle/shishua.h for the
provenance caveat).  Values are stored as IEEE-754 hex so nothing is lost."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import oracle_api as oa  # noqa: E402
from cases import PRICE_CASES, RNG_CASES  # noqa: E402


def hexlist(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    assert oa.have_ref(), "build oracle/_ref first: make -C oracle"
    out = {"prices": {}, "rng": {}, "shishua": {}}
    for name, (payoff, T, K, steps, params, n) in PRICE_CASES.items():
        c = oa.Contract(payoff, T, K, steps, params)
        entry = {}
        for mode, label in ((oa.NORMAL_F32, "f32"), (oa.NORMAL_F64, "f64")):
            entry[label] = hexlist(c.ref_price(n, threads=1, normal_mode=mode))
        out["prices"][name] = entry
        print(name, [float.fromhex(x) for x in entry["f32"]][:4])
    for name, (size, seed, pat, reps) in RNG_CASES.items():
        kinds = [1 if ch == "u" else 0 for ch in pat] * reps
        entry = {}
        for mode, label in ((oa.NORMAL_F32, "f32"), (oa.NORMAL_F64, "f64")):
            seq = oa.ref_rng_sequence(size, seed, kinds, mode)
            entry[label] = hexlist(seq)
        out["rng"][name] = entry
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(out, f)
    print("wrote reference_outputs.json")


if __name__ == "__main__":
    main()
