"""Geometric-Asian control variate (the control the reference itself names at
src/inc/HSimulation.h:51).

Next to the arithmetic average A = sum w_i X_i / T of AAsianCallNonAdaptive
(src/inc/AsianContract.h:25-34) the path kernel accumulates Y = sum w_i ln X_i / sum w_i with the
same weights and uses c_j = max(exp(Y) - K_j, 0) as the control of option j.  Its mean is the
price of a discretely monitored geometric-Asian call under Heston (r = 0), computed on the host
(csrc/geo_asian_host.cu).  Checked here:
  * the host value against an independent numpy restatement (other inversion formula, other
    quadrature) and against the oracle's Monte-Carlo mean of the same control;
  * the estimator (hexo_gpu_finish) on oracle sums: price within the plain Monte-Carlo error,
    standard error >= 5 x smaller at the money on cfg1's shape;
  * on the GPU: all five sums equal the oracle's on the same streams, and the same two
    properties at 1e6 paths.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa
import hestonexotics_b200 as hx
from hestonexotics_b200 import pricing

ASIAN = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
P0 = hx.HParams(*oa.DEFAULT_PARAMS)


def chains_of(T, K):
    return [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]


# ---- independent restatement of E max(G - K, 0) -----------------------------------------------

def log_weights(expiries, steps, k, exact):
    """normalised weights of Y on ln X_0..ln X_N for maturity k and the step widths, from the
    public schedule (what the Asian policy applies: AsianContract.h:25-34, HSimulation.tpp:42-44)"""
    sched = hx.schedule(expiries, steps, "exact" if exact else "reference")
    N = sum(s[0] for s in sched[:k + 1])
    w, h = np.zeros(N + 1), np.zeros(N + 1)
    cur = prev = 0
    for s, (n, hs_, wk, T) in enumerate(sched[:k + 1]):
        if n == 0:
            continue
        if s > 0:
            carry = (sched[s - 1][1] if exact else hs_) / 2
            w[cur] += carry
            w[prev] += carry
        a, b = cur, cur + n
        h[a + 1:b + 1] = hs_
        w[a] += hs_ / 2
        w[b - 1] -= hs_ / 2
        w[a + 1:b] += hs_
        cur, prev = b, b - 1
    n, hk, wk, T = sched[k]
    if exact:
        w[cur] += hk / 2
        w[prev] += hk / 2
    else:
        w[cur] += wk
        w[prev] -= wk
    return w / w.sum(), h


def log_mgf(z, params, S, w, h):
    """ln E exp(z Y), vectorised over z: backward recursion of the one-step affine transform"""
    v0, th, rho, ka, si = params
    z = np.asarray(z, dtype=complex)
    om = np.cumsum(w[::-1])[::-1]
    b = np.zeros_like(z)
    acc = z * om[0] * np.log(S)
    for m in range(len(w) - 1, 0, -1):
        a = z * om[m]
        beta = ka - rho * si * a
        D = np.sqrt(beta * beta - si * si * (a * a - a))
        Bp, Bm = (beta + D) / si ** 2, (beta - D) / si ** 2
        y0 = (b - Bm) / (b - Bp)
        y = y0 * np.exp(-D * h[m])
        acc = acc + ka * th * (Bm * h[m] - 2 / si ** 2 * np.log((1 - y) / (1 - y0)))
        b = (Bm - Bp * y) / (1 - y)
    return acc + b * v0


def geometric_call_gil_pelaez(params, S, w, h, strikes):
    """E max(e^Y - K, 0) = F P1 - K P2 with the two Gil-Pelaez probabilities (NOT the Lewis form
    the product uses), 64-point Gauss-Legendre on the panels [0,1], [1,4], [4,16], ... (the
    integrands Im[e^{-iuk} phi(u)] / u are finite at u = 0)."""
    x, gw = np.polynomial.legendre.leggauss(64)
    edges = [0.0, 1.0, 4.0, 16.0, 64.0, 256.0, 1024.0]
    u = np.concatenate([0.5 * (a + b) + 0.5 * (b - a) * x for a, b in zip(edges, edges[1:])])
    wt = np.concatenate([0.5 * (b - a) * gw for a, b in zip(edges, edges[1:])])
    lnF = log_mgf(np.array([1.0]), params, S, w, h)[0].real
    phi2 = np.exp(log_mgf(1j * u, params, S, w, h))
    phi1 = np.exp(log_mgf(1j * u + 1.0, params, S, w, h) - lnF)
    out = []
    for K in strikes:
        k = np.log(K)
        p2 = 0.5 + (wt * (np.exp(-1j * u * k) * phi2 / (1j * u)).real).sum() / np.pi
        p1 = 0.5 + (wt * (np.exp(-1j * u * k) * phi1 / (1j * u)).real).sum() / np.pi
        out.append(np.exp(lnF) * p1 - K * p2)
    return np.array(out)


GEO_CASES = [
    ("cfg1_shape", [1.0], [[80.0, 100.0, 120.0]], 252, oa.DEFAULT_PARAMS, False),
    ("quirk_64", [1.0], [[90.0, 100.0]], 64, oa.DEFAULT_PARAMS, False),
    ("two_maturities", [0.3, 1.0], [[95.0, 100.0], [100.0, 110.0]], 20, oa.DEFAULT_PARAMS, False),
    ("exact_grid", [0.25, 0.5], [[100.0], [100.0]], 16, oa.DEFAULT_PARAMS, True),
    ("stiff", [2.0], [[100.0]], 100, oa.STIFF_PARAMS, False),
]


@pytest.mark.parametrize("name,T,K,steps,params,exact", GEO_CASES, ids=[c[0] for c in GEO_CASES])
def test_host_geometric_mean_vs_independent_restatement(hexo_lib, name, T, K, steps, params, exact):
    got = hx.geometric_asian_means(hx.HParams(*params), 100.0, chains_of(T, K), steps,
                                   "exact" if exact else "reference")
    want = []
    for k in range(len(T)):
        w, h = log_weights(T, steps, k, exact)
        assert abs(w.sum() - 1.0) < 1e-12
        want += list(geometric_call_gil_pelaez(params, 100.0, w, h, K[k]))
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-9)


def test_host_geometric_mean_strike_zero_is_the_forward(hexo_lib):
    """K = 0: the control is G itself, E[G] = exp(log-mgf(1)) < S (Jensen)."""
    got = hx.geometric_asian_means(P0, 100.0, chains_of([1.0], [[0.0, 1e-9]]), 50)
    w, h = log_weights([1.0], 50, 0, False)
    F = np.exp(log_mgf(np.array([1.0]), oa.DEFAULT_PARAMS, 100.0, w, h)[0].real)
    assert abs(got[0] - F) < 1e-12 and abs(got[1] - F) < 1e-6 and 98.0 < F < 100.0


@pytest.mark.parametrize("name,T,K,steps,params,exact", GEO_CASES[:4], ids=[c[0] for c in GEO_CASES[:4]])
def test_host_geometric_mean_vs_oracle_monte_carlo(hexo_lib, name, T, K, steps, params, exact):
    """The simulated mean of the control (oracle, QE scheme) agrees with the semi-analytic mean
    of the exact process within Monte-Carlo error (+ the scheme's small weak error)."""
    c = oa.Contract(oa.ASIAN, T, K, steps, params)
    n = 150_000
    s = c.price_stream_geo(5, n, 128, normal_mode=oa.NORMAL_F64, exact_grid=exact)
    m = c.n_opts
    mc = s[3 * m:4 * m] / n
    se = np.sqrt((s[4 * m:] / n - mc ** 2) / n)
    eg = hx.geometric_asian_means(hx.HParams(*params), 100.0, chains_of(T, K), steps,
                                  "exact" if exact else "reference")
    assert np.all(np.abs(mc - eg) < 4 * se + 2e-3 * eg + 1e-4), (mc, eg, se)


def geo_estimate(sums, n, eg):
    m = len(eg)
    sp, sq, sx, sc, sc2 = (sums[i * m:(i + 1) * m] for i in range(5))
    mean, mc = sp / n, sc / n
    var = (sq - n * mean ** 2) / (n - 1)
    vc = (sc2 - n * mc ** 2) / (n - 1)
    cov = (sx - n * mean * mc) / (n - 1)
    beta = cov / vc
    return mean - beta * (mc - eg), np.sqrt(np.maximum(var - beta * cov, 0) / n)


def test_finish_with_geometric_control_on_oracle_sums(hexo_lib):
    """cfg1's shape (Asian, T = 1, 252 steps): hexo_gpu_finish on the oracle's sums = the textbook
    estimator; the price stays within the plain Monte-Carlo error and the standard error shrinks
    by more than 5 x at the money (the verdict's bar), 10 x in the money."""
    T, K, steps, n = [1.0], [[80.0, 100.0, 110.0]], 252, 60_000
    c = oa.Contract(oa.ASIAN, T, K, steps)
    sums = c.price_stream_geo(9, n, 128, normal_mode=oa.NORMAL_F64)
    rq = pricing._Request(ASIAN, P0, 100.0, chains_of(T, K), n, None, steps, 9, "f64", 128,
                          control_variate="geometric")
    assert hexo_lib.hexo_gpu_sums_len(C.byref(rq.req)) == sums.size == 15
    price, se = pricing._finish(rq, sums)
    eg = hx.geometric_asian_means(P0, 100.0, chains_of(T, K), steps)
    wp, wse = geo_estimate(sums, n, eg)
    np.testing.assert_allclose(price, wp, rtol=1e-12)
    np.testing.assert_allclose(se, wse, rtol=1e-9)
    plain = pricing._Request(ASIAN, P0, 100.0, chains_of(T, K), n, None, steps, 9, "f64", 128)
    p0, se0 = pricing._finish(plain, sums[:6])
    assert np.all(np.abs(price - p0) < 4 * se0 + 0.01)
    ratio = se0 / se
    assert ratio[1] > 5 and ratio[0] > 10 and ratio[2] > 3, ratio


def test_geometric_control_is_refused_for_european(hexo_lib):
    with pytest.raises(ValueError):
        pricing._Request(ASIAN, P0, 100.0, chains_of([1.0], [[100.0]]), 10, None, 10, 1, "f32", 1,
                         control_variate="harmonic")
    rq = pricing._Request(hx.HQEAnderson(hx.EuropeanCallNonAdaptive), P0, 100.0,
                          chains_of([1.0], [[100.0]]), 10, None, 10, 1, "f32", 1,
                          control_variate="geometric")
    out = np.zeros(1)
    assert hexo_lib.hexo_gpu_finish(C.byref(rq.req), np.zeros(5).ctypes.data_as(
        pricing._lib.c_double_p), out.ctypes.data_as(pricing._lib.c_double_p), None) == -1


# ------------------------------------------------------------------------------------------ GPU
GPU_CASES = [
    ("asian_chain", [1.0], [[90.0, 100.0, 110.0]], 64, 3000, 96, False, "reference"),
    ("two_maturities_exact", [0.25, 1.0], [[95.0, 105.0], [100.0]], 20, 1501, 77, True, "reference"),
    ("three_maturities", [0.3, 0.6, 2.0], [[100.0], [100.0], [90.0, 120.0]], 40, 2000, 70, False,
     "reference"),
    ("martingale_drift", [1.0], [[100.0, 105.0]], 48, 2000, 64, False, "martingale"),
    ("70_strikes", [0.5], [list(np.linspace(70.0, 130.0, 70))], 16, 900, 64, False, "reference"),
    ("one_path", [1.0], [[100.0]], 16, 1, 1, False, "reference"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,T,K,steps,n_paths,n_streams,exact,drift", GPU_CASES,
                         ids=[c[0] for c in GPU_CASES])
def test_gpu_geometric_sums_match_oracle(gpu, name, T, K, steps, n_paths, n_streams, exact, drift):
    c = oa.Contract(oa.ASIAN, T, K, steps, drift_mode=1 if drift == "martingale" else 0)
    want = c.price_stream_geo(13, n_paths, n_streams, normal_mode=oa.NORMAL_F64, exact_grid=exact)
    kw = dict(seed=13, normal_mode="f64", n_streams=n_streams, drift=drift,
              time_grid="exact" if exact else "reference")
    res = hx.price_full(ASIAN, P0, 100.0, chains_of(T, K), n_paths, None, steps,
                        control_variate="geometric", **kw)
    assert res.sums.size == want.size == 5 * c.n_opts
    assert np.all(np.abs(res.sums - want) <= 1e-10 * np.abs(want) + 1e-9 * n_paths)
    plain = hx.price_full(ASIAN, P0, 100.0, chains_of(T, K), n_paths, None, steps, **kw)
    # the plain sums are the plain kernel's (another instantiation: equal up to the last bits of
    # the compiler's multiply-add contraction)
    np.testing.assert_allclose(res.sums[:2 * c.n_opts], plain.sums, rtol=1e-12)


@pytest.mark.gpu
def test_gpu_geometric_control_reduces_the_error_and_keeps_the_price(gpu):
    """cfg1's shape, 1e6 paths x 252 steps: standard errors >= 5 x smaller at the money (>= 10 x
    in the money), prices within the plain Monte-Carlo error; the same through hexo_gpu_price_multi
    and hexo_gpu_price_batch; also with as-built single-precision normals."""
    Ks = [80.0, 90.0, 100.0, 110.0]
    ch = chains_of([1.0], [Ks])
    for mode in ("f32", "f64"):
        plain = hx.price_full(ASIAN, P0, 100.0, ch, 1_000_000, 4, 252, seed=3, normal_mode=mode)
        cv = hx.price_full(ASIAN, P0, 100.0, ch, 1_000_000, 4, 252, seed=3, normal_mode=mode,
                           control_variate="geometric")
        ratio = plain.stderr / cv.stderr
        print(f"geometric control, {mode}: plain {plain.prices} +- {plain.stderr}, "
              f"cv {cv.prices} +- {cv.stderr}, ratio {ratio}")
        assert ratio[2] > 5 and ratio[0] > 10 and ratio[1] > 10 and ratio[3] > 3, ratio
        assert np.all(np.abs(cv.prices - plain.prices) < 4 * plain.stderr + 0.005)
    pm, sem = hx.price_multi(ASIAN, P0, 100.0, ch, 1_000_000, 4, 252, n_gpus=1, seed=3,
                             normal_mode="f64", n_streams=cv.n_streams, control_variate="geometric")
    np.testing.assert_allclose(pm, cv.prices, rtol=1e-12)
    pb, seb, _ = hx.price_batch(ASIAN, [P0, P0], 100.0, ch, 1_000_000, 4, 252, seeds=3,
                                normal_mode="f64", n_streams=cv.n_streams,
                                control_variate="geometric")
    assert np.array_equal(pb[0], cv.prices) and np.array_equal(seb[1], cv.stderr)


@pytest.mark.gpu
def test_gpu_geometric_control_cfg4_grid(gpu):
    """1024 steps (the reference grid's last-step rule active, weights sum to T (1 - 1/1024)):
    the normalised geometric control still tracks the arithmetic payoff (correlation > 0.999)."""
    ch = chains_of([1.0], [[100.0]])
    plain = hx.price_full(ASIAN, P0, 100.0, ch, 2_000_000, 1, 1024, seed=5)
    cv = hx.price_full(ASIAN, P0, 100.0, ch, 2_000_000, 1, 1024, seed=5, control_variate="geometric")
    assert plain.stderr[0] / cv.stderr[0] > 20
    assert abs(cv.prices[0] - plain.prices[0]) < 4 * plain.stderr[0] + 0.005


def _random_geo_contract(rng):
    n_chains = int(rng.integers(1, 4))
    T = np.cumsum(rng.uniform(0.05, 0.9, size=n_chains)).tolist()
    S = float(rng.uniform(30.0, 300.0))
    K = [sorted((S * rng.uniform(0.7, 1.3, size=int(rng.integers(1, 5)))).tolist())
         for _ in range(n_chains)]
    steps = int(rng.integers(2, 70))
    params = (float(rng.uniform(0.01, 0.2)), float(rng.uniform(0.01, 0.2)),
              float(rng.uniform(-0.95, 0.3)), float(rng.uniform(0.3, 6.0)),
              float(rng.uniform(0.1, 1.2)))
    n_paths = int(rng.integers(1, 1200))
    return T, K, steps, params, S, n_paths, int(rng.integers(1, min(n_paths, 200) + 1))


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(10))
def test_gpu_geometric_sums_random_contracts(gpu, i):
    """Seeded random Asian contracts (maturities, ragged chains, step counts, parameters, spot,
    path / stream counts, both grids, both drifts): all five sums equal the oracle's, and the
    host mean of the control is finite and below the forward."""
    rng = np.random.default_rng(4000 + i)
    T, K, steps, params, S, n_paths, n_streams = _random_geo_contract(rng)
    exact, mart = bool(i & 1), bool(i & 2)
    seed = int(rng.integers(0, 2 ** 62))
    c = oa.Contract(oa.ASIAN, T, K, steps, params, S, drift_mode=int(mart))
    want = c.price_stream_geo(seed, n_paths, n_streams, normal_mode=oa.NORMAL_F64, exact_grid=exact)
    res = hx.price_full(ASIAN, hx.HParams(*params), S, chains_of(T, K), n_paths, None, steps,
                        seed=seed, normal_mode="f64", n_streams=n_streams,
                        time_grid="exact" if exact else "reference",
                        drift="martingale" if mart else "reference", control_variate="geometric")
    m = c.n_opts
    floor = 1e-12 * n_paths * S
    scale = np.concatenate([np.abs(want[:m]), np.abs(want[m:2 * m]), np.abs(want[m:2 * m]),
                            np.abs(want[3 * m:4 * m]), np.abs(want[4 * m:])])
    assert np.all(np.abs(res.sums - want) <= 1e-10 * scale + floor * np.repeat([1, S, S, 1, S], m))
    eg = hx.geometric_asian_means(hx.HParams(*params), S, chains_of(T, K), steps,
                                  "exact" if exact else "reference")
    assert np.all(np.isfinite(eg)) and np.all(eg >= 0) and np.all(eg < S)
    assert np.all(np.isfinite(res.prices)) and np.all(res.stderr >= 0)


@pytest.mark.gpu
def test_gpu_geometric_control_against_the_full_size_plain_estimate(gpu):
    """Size-independent check on cfg4's grid: 10^7 paths WITH the control against 10^9 paths
    without (the BASELINE job, 6 s).  The two estimators converge to limits that differ by the
    part of the QE scheme's weak error the control removes; measured difference and errors are
    printed, the bound is 5e-3 (0.1 % of the price)."""
    ch = chains_of([1.0], [[100.0]])
    plain = hx.price_full(ASIAN, P0, 100.0, ch, 1_000_000_000, 1, 1024, seed=1)
    cv = hx.price_full(ASIAN, P0, 100.0, ch, 10_000_000, 1, 1024, seed=2, control_variate="geometric")
    d = cv.prices[0] - plain.prices[0]
    print(f"cfg4 grid: plain 1e9 paths {plain.prices[0]:.6f} +- {plain.stderr[0]:.6f}; geometric control "
          f"1e7 paths {cv.prices[0]:.6f} +- {cv.stderr[0]:.6f}; difference {d:+.6f}")
    assert cv.stderr[0] < 0.5 * plain.stderr[0]          # 100 x fewer paths, smaller error
    assert abs(d) < 5e-3
