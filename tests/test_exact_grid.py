"""Corrected time grid (`schedule_mode = HEXO_SCHEDULE_EXACT`): the grid ends on
every expiry, the Asian average is the full trapezoid rule.  Not the reference's behaviour -- the
oracle holds an independent restatement of the definition (oracle_price_stream_exact) and the
closed-form European price anchors it."""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa


def test_exact_schedule_host(hexo_lib):
    import hestonexotics_b200 as hx
    # single maturity: exactly `steps` steps of T/steps, last step counts fully
    for T, steps in [(1.0, 252), (1.0, 1024), (0.37, 100), (10.0, 2520), (1.0, 1)]:
        (n, h, w, ex), = hx.schedule([T], steps, time_grid="exact")
        assert (n, w, ex) == (steps, 1.0, T) and h == T / steps
    # the reference's grid for comparison: 253 calls at (1, 252)
    assert hx.schedule([1.0], 252)[0][0] == 253
    # several maturities: every segment ends on its expiry
    T = [0.25, 0.5, 1.0, 1.7]
    seg = hx.schedule(T, 100, time_grid="exact")
    prev = 0.0
    for (n, h, w, ex), t in zip(seg, T):
        assert ex == t and w == 1.0 and n >= 1
        assert n == max(1, round((t - prev) * 100 / t))
        assert abs(n * h - (t - prev)) < 1e-15
        prev = t
    assert [s[0] for s in seg] == [100, 50, 50, 41]


def test_exact_schedule_rejects_bad_expiries(hexo_lib):
    seg = (hexo_lib_segment() * 2)()
    ex = np.array([1.0, 0.5])
    rc = hexo_lib.hexo_gpu_schedule_exact(ex.ctypes.data_as(C.POINTER(C.c_double)), 2, 10, seg)
    assert rc == -2


def hexo_lib_segment():
    from hestonexotics_b200 import _lib
    return _lib.HexoSegment


def test_oracle_exact_grid_removes_the_last_trapezoid_bias():
    """Reference grid: the Asian average at steps=64 (lands on T: last trapezoid replaced) sits
    about S/steps below the one at steps=63 (falls short of T: one more step, full trapezoid)
    -- SURVEY finding 6.  On the exact grid neighbouring step counts agree."""
    n, ns = 20000, 100
    ref = {st: oa.Contract(oa.ASIAN, [1.0], [[0.0]], st).price_stream(5, n, ns,
                                                                       normal_mode=oa.NORMAL_F64)[0][0] / n
           for st in (64,)}
    exa = {st: oa.Contract(oa.ASIAN, [1.0], [[0.0]], st).price_stream(
               5, n, ns, normal_mode=oa.NORMAL_F64, exact_grid=True)[0][0] / n for st in (64,)}
    # strike 0: the payoff IS the average; E[average] = S = 100 on the exact grid
    assert abs(exa[64] - 100.0) < 0.35
    # reference grid, steps = 64 (1/64 is exact in binary: the grid lands on T): average short
    # by about h/2 * (X_N + X_{N-1}) - (X_N - X_{N-1}) ~ S/steps
    assert 1.2 < exa[64] - ref[64] < 1.9


def test_oracle_exact_grid_european_matches_closed_form():
    from heston_cf import heston_call
    c = oa.Contract(oa.EUROPEAN, [0.5, 1.0], [[100.0], [90.0, 110.0]], 64)
    n = 60000
    sm, sq = c.price_stream(9, n, 128, normal_mode=oa.NORMAL_F64, exact_grid=True)
    se = np.sqrt((sq / n - (sm / n) ** 2) / n)
    want = np.array([heston_call(100.0, 100.0, 0.5, *oa.DEFAULT_PARAMS),
                     heston_call(100.0, 90.0, 1.0, *oa.DEFAULT_PARAMS),
                     heston_call(100.0, 110.0, 1.0, *oa.DEFAULT_PARAMS)])
    assert np.all(np.abs(sm / n - want) < 4 * se + 0.03)


# --------------------------------------------------------------------------- GPU
EXACT_CASES = [
    ("asian_1024", oa.ASIAN, [1.0], [[100.0]], 1024, 300, 64),
    ("asian_252", oa.ASIAN, [1.0], [[90.0, 100.0, 110.0]], 252, 700, 96),
    ("asian_four_maturities", oa.ASIAN, [0.25, 0.5, 1.0, 1.7], [[100.0], [95.0, 105.0], [100.0], [80.0]],
     60, 1500, 130),
    ("euro_three_maturities", oa.EUROPEAN, [0.3, 0.31, 2.0], [[100.0], [100.0], [90.0, 120.0]], 40,
     2000, 77),
    ("euro_one_step", oa.EUROPEAN, [1.0], [[100.0]], 1, 3000, 64),
    ("asian_one_step", oa.ASIAN, [0.5], [[100.0]], 1, 3000, 64),
    ("asian_close_maturities", oa.ASIAN, [1.0, 1.001, 1.002], [[100.0]] * 3, 50, 1200, 64),
]


@pytest.mark.gpu
@pytest.mark.parametrize("rng_mode", [0, 1])
@pytest.mark.parametrize("name,payoff,T,K,steps,n_paths,n_streams", EXACT_CASES,
                         ids=[c[0] for c in EXACT_CASES])
def test_gpu_exact_grid_sums_match_oracle(gpu, name, payoff, T, K, steps, n_paths, n_streams,
                                          rng_mode):
    import hestonexotics_b200 as hx
    c = oa.Contract(payoff, T, K, steps)
    sm, sq = c.price_stream(21, n_paths, n_streams, normal_mode=oa.NORMAL_F64, rng_mode=rng_mode,
                            exact_grid=True)
    pol = hx.AAsianCallNonAdaptive if payoff == oa.ASIAN else hx.EuropeanCallNonAdaptive
    chains = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
    res = hx.price_full(hx.HQEAnderson(pol), hx.HParams(*oa.DEFAULT_PARAMS), 100.0, chains, n_paths,
                        None, steps, seed=21, normal_mode="f64", n_streams=n_streams,
                        rng=("shishua", "philox")[rng_mode], time_grid="exact")
    n = c.n_opts
    floor = 1e-12 * n_paths * 100.0
    assert np.all(np.abs(res.sums[:n] - sm) <= 1e-10 * np.abs(sm) + floor)
    assert np.all(np.abs(res.sums[n:] - sq) <= 2e-10 * np.abs(sq) + 100 * floor)
    want_steps = sum(s[0] for s in hx.schedule(T, steps, time_grid="exact"))
    assert res.steps_per_path == want_steps


@pytest.mark.gpu
def test_gpu_exact_grid_asian_average_is_unbiased(gpu):
    """Strike 0: the payoff is the average itself, whose expectation is S.  On the reference grid
    at steps = 1024 it is short by about S/steps (finding 6); on the exact grid it is not."""
    import hestonexotics_b200 as hx
    p = hx.HParams(*oa.DEFAULT_PARAMS)
    ch = [hx.OptionsChain.from_strikes(1.0, [0.0])]
    A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
    ref = hx.price_full(A, p, 100.0, ch, 4_000_000, 1, 1024, seed=2)
    exa = hx.price_full(A, p, 100.0, ch, 4_000_000, 1, 1024, seed=2, time_grid="exact")
    assert abs(exa.prices[0] - 100.0) < 5 * exa.stderr[0] + 0.01   # + QE drift bias
    assert abs((exa.prices[0] - ref.prices[0]) - 100.0 / 1024) < 0.01


@pytest.mark.gpu
def test_gpu_unknown_schedule_mode_is_refused(gpu):
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import _lib, pricing
    rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                          100.0, [hx.OptionsChain.from_strikes(1.0, [100.0])], 100, 1, 10, 1,
                          "f32", 32)
    rq.req.schedule_mode = 3
    sums = np.zeros(2)
    assert gpu.hexo_gpu_price_shard(C.byref(rq.req), 0, 32, sums.ctypes.data_as(_lib.c_double_p),
                                    None) == -1
