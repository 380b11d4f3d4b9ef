"""The drop-in from the reference's own language: oracle/_ref/hexo_dropin is compiled from the
reference's unchanged C++ headers plus hestonexotics_b200/cpp/hexo_gpu_adapter.hpp and calls
HSimulation::price<Scheme> (CPU, reference code) and HSimulation::price_gpu<Scheme> (C ABI -> CUDA)
on the same synthetic option chain, like the PRICE case of src/Main.cpp:75-96."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "hexo_dropin")


@pytest.mark.parametrize("kind", ["asian", "european"])
def test_cpp_dropin_matches_reference_cpu(gpu, kind):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/hexo_dropin not built (needs the reference tree at build time)")
    n_cpu, n_gpu = 100000, 4_000_000
    out = subprocess.run([BIN, kind, str(n_cpu), str(n_gpu), "100"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout)
    ref, got, se = (np.array(res[k]) for k in ("reference", "gpu", "gpu_stderr"))
    assert ref.shape == got.shape == (9,)
    se_ref = se * np.sqrt(n_gpu / n_cpu)           # same payoff variance, fewer paths
    z = (got - ref) / np.hypot(se, se_ref)
    assert np.abs(z).max() < 3.5, z                # 9 options: 3.5 sigma ~ 0.4 % false alarm
    # chain-major order: within each maturity prices fall with the strike (90, 100, 110)
    g = got.reshape(3, 3)
    assert (np.diff(g, axis=1) < 0).all() and (np.diff(g[:, 1]) > 0).all()


def test_cli_demo_prints_reference_format(gpu):
    """Row f2: `hexo -p asian all <SYM>` (src/Main.cpp:75-96) with a synthetic chain file instead
    of the Tradier download and fixed HParams instead of the SQLite DB; pricing through
    price_gpu<>, output loop in the reference's own format."""
    exe = os.path.join(ROOT, "oracle", "_ref", "hexo_cli_demo")
    chain = os.path.join(ROOT, "tests", "golden", "synthetic_chain.csv")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/hexo_cli_demo not built (needs the reference tree at build time)")
    out = subprocess.run([exe, chain], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("params, v0: 0.04\tv_m: 0.04\trho: -0.7\tkappa: 2\tsigma: 0.5")
    body = lines[1:]
    assert len(body) == 9
    prices = []
    for ln in body:
        fields = dict(f.split(": ") for f in ln.split("\t"))
        assert list(fields) == ["S", "strike", "bid", "ask", "asian-option-price", "volume",
                                "imp vol", "lb", "expiry time"]
        prices.append(float(fields["asian-option-price"]))
        assert float(fields["imp vol"]) > 0.05
    pr = np.array(prices).reshape(3, 3)
    assert (np.diff(pr, axis=1) < 0).all()          # falls with the strike within each expiry
    assert 1.0 < pr[0, 1] < 2.0 and pr[2, 1] > pr[1, 1] > pr[0, 1]   # ATM Asian grows with expiry


def test_cli_demo_with_the_geometric_control(gpu):
    """The same CLI with `geometric` appended: GpuPriceOptions::control_variate through the C++
    adapter.  10^5 paths with the control agree with 4*10^6 plain paths within a few cents."""
    exe = os.path.join(ROOT, "oracle", "_ref", "hexo_cli_demo")
    chain = os.path.join(ROOT, "tests", "golden", "synthetic_chain.csv")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/hexo_cli_demo not built (needs the reference tree at build time)")

    def run(*extra):
        out = subprocess.run([exe, chain, "0.04", "0.04", "-0.7", "2.0", "0.5", *extra],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        return np.array([float(dict(f.split(": ") for f in ln.split("\t"))["asian-option-price"])
                         for ln in out.stdout.strip().splitlines()[1:]])
    plain = run("4000000", "252")
    cv = run("100000", "252", "geometric")
    assert plain.shape == cv.shape == (9,)
    assert np.abs(cv - plain).max() < 0.02, (cv, plain)
