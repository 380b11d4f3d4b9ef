"""Control variate (the reference only suggests one, src/inc/HSimulation.h:51):
c = final value - S (arithmetic average or terminal spot minus the initial spot).  The simulated
model has no drift, so E[X_j] = S and E[c] follows from the weights the payoff policy applies:
0 for the European payoff and for the full trapezoid rule, S (W/T - 1) on the reference's grid,
whose last trapezoid is replaced when the steps land on the expiry (SURVEY finding 6).  The
kernel accumulates [sum pf c] per option and [sum c | sum c^2] per maturity next to the plain
sums; hexo_gpu_finish turns them into price = mean(pf) - beta (mean(c) - E[c])."""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa


def control_means(payoff, S, expiries, steps, exact=False):
    """E[c] per maturity, restated in numpy from the step schedule: the sum of the trapezoid
    weights the Asian policy actually applies (AsianContract.h:25-34, HSimulation.tpp:42-44)."""
    import hestonexotics_b200 as hx
    if payoff != oa.ASIAN:
        return np.zeros(len(expiries))
    sched = hx.schedule(expiries, steps, "exact" if exact else "reference")
    out, weight = [], 0.0
    for k, (n, h, w, T) in enumerate(sched):
        if n > 0:
            if k > 0:   # the trapezoid of the step that crossed the previous expiry
                weight += sched[k - 1][1] if exact else h
            weight += h * (n - 1)
        out.append(S * ((weight + (h if exact else 0.0)) / T - 1.0))
    return np.array(out)


def cv_estimate(sums, n_paths, offsets, ec=None):
    """Reference implementation of the estimator in numpy (what hexo_gpu_finish must compute);
    ec = known mean of the control per maturity."""
    n_opts, n_ch = int(offsets[-1]), len(offsets) - 1
    ec = np.zeros(n_ch) if ec is None else ec
    sp, sq, sx = sums[:n_opts], sums[n_opts:2 * n_opts], sums[2 * n_opts:3 * n_opts]
    sc, sc2 = sums[3 * n_opts:3 * n_opts + n_ch], sums[3 * n_opts + n_ch:]
    n = float(n_paths)
    prices, se = np.zeros(n_opts), np.zeros(n_opts)
    for k in range(n_ch):
        mc = sc[k] / n
        vc = (sc2[k] - n * mc * mc) / (n - 1)
        for j in range(offsets[k], offsets[k + 1]):
            m = sp[j] / n
            var = (sq[j] - n * m * m) / (n - 1)
            cov = (sx[j] - n * m * mc) / (n - 1)
            beta = cov / vc
            prices[j] = m - beta * (mc - ec[k])
            se[j] = np.sqrt(max(0.0, var - beta * cov) / n)
    return prices, se


def test_finish_matches_numpy_estimator(hexo_lib):
    """Host only: hexo_gpu_finish on oracle sums = the textbook control-variate estimator."""
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import pricing
    T, K = [0.5, 1.0], [[90.0, 100.0], [100.0, 110.0, 120.0]]
    c = oa.Contract(oa.ASIAN, T, K, 30)
    n = 20000
    sums = c.price_stream_cv(4, n, 64, normal_mode=oa.NORMAL_F64)
    rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                          100.0, [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)], n,
                          None, 30, 4, "f64", 64, control_variate="underlying")
    assert hexo_lib.hexo_gpu_sums_len(C.byref(rq.req)) == sums.size == 3 * 5 + 2 * 2
    prices, se = pricing._finish(rq, sums)
    want_p, want_se = cv_estimate(sums, n, c.offsets, control_means(oa.ASIAN, 100.0, T, 30))
    np.testing.assert_allclose(prices, want_p, rtol=1e-12)
    np.testing.assert_allclose(se, want_se, rtol=1e-9)
    # and it is a variance reduction: compare with the plain standard error
    plain = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                             100.0, [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)], n,
                             None, 30, 4, "f64", 64)
    p0, se0 = pricing._finish(plain, sums[:10])
    assert np.all(se <= se0) and np.all(se[:3] < 0.75 * se0[:3])   # little to gain far out of the money
    assert np.all(np.abs(prices - p0) < 4 * se0)
    np.testing.assert_allclose(p0, sums[:5] / n, rtol=1e-15)


@pytest.mark.parametrize("steps,exact", [(64, False), (252, False), (64, True)])
def test_control_mean_on_the_reference_grid(hexo_lib, steps, exact):
    """on a power-of-two step count the reference's grid replaces the last trapezoid,
    E[average] = S (1 - 1/steps), so the control has mean -S/steps (-1.5625 at 64 steps), not 0.
    The estimator must subtract the known mean: the simulated mean of c agrees with it and the
    control-variate price agrees with the plain one (host only: oracle sums + hexo_gpu_finish)."""
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import pricing
    T, K, n = [1.0], [[80.0, 100.0]], 200_000
    c = oa.Contract(oa.ASIAN, T, K, steps)
    sums = c.price_stream_cv(21, n, 256, normal_mode=oa.NORMAL_F64, exact_grid=exact)
    ec = control_means(oa.ASIAN, 100.0, T, steps, exact)
    if not exact and steps == 64:
        assert abs(ec[0] + 100.0 / 64) < 1e-12
    if exact:
        assert abs(ec[0]) < 1e-12
    mean_c = sums[3 * 2] / n
    se_c = np.sqrt((sums[3 * 2 + 1] / n - mean_c ** 2) / n)
    assert abs(mean_c - ec[0]) < 4 * se_c + 0.01          # + the QE drift error of E[X_j] = S
    chains = [hx.OptionsChain.from_strikes(1.0, K[0])]
    grid = "exact" if exact else "reference"
    rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                          100.0, chains, n, None, steps, 21, "f64", 256, time_grid=grid,
                          control_variate="underlying")
    plain = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                             100.0, chains, n, None, steps, 21, "f64", 256, time_grid=grid)
    p_cv, se_cv = pricing._finish(rq, sums)
    p0, se0 = pricing._finish(plain, sums[:4])
    assert np.all(np.abs(p_cv - p0) < 4 * se0), (p_cv, p0, se0)   # was 64 standard errors apart
    assert se_cv[0] < 0.3 * se0[0]


def test_oracle_cv_sums_extend_the_plain_sums():
    c = oa.Contract(oa.EUROPEAN, [0.25, 1.0], [[95.0], [100.0, 105.0]], 20)
    sm, sq = c.price_stream(5, 1003, 17, normal_mode=oa.NORMAL_F64)
    full = c.price_stream_cv(5, 1003, 17, normal_mode=oa.NORMAL_F64)
    np.testing.assert_array_equal(full[:3], sm)
    np.testing.assert_array_equal(full[3:6], sq)
    a = c.price_stream_cv(5, 1003, 17, 0, 9, normal_mode=oa.NORMAL_F64)
    b = c.price_stream_cv(5, 1003, 17, 9, 8, normal_mode=oa.NORMAL_F64)
    np.testing.assert_allclose(a + b, full, rtol=1e-12, atol=1e-9)


def test_distributed_cv_with_gloo_matches_single_process():
    """world_size 2 over gloo: shards of the control-variate sums add up to the single-process
    sums and every rank gets the same control-variate price."""
    import torch.multiprocessing as mp
    mp.spawn(_cv_worker, args=(2,), nprocs=2, join=True)


def _cv_worker(rank, world):
    import os
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    import oracle_api as oa
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import pricing
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29541", rank=rank, world_size=world)
    try:
        T, K, steps, n = [0.5, 1.0], [[100.0], [90.0, 110.0]], 12, 4001
        c = oa.Contract(oa.ASIAN, T, K, steps)

        import torch

        def shard(rq, begin, count, world_, stats):   # the oracle stands in for the GPU shard
            return torch.from_numpy(c.price_stream_cv(
                int(rq.req.seed), int(rq.req.n_paths), int(rq.req.n_streams), begin, count,
                normal_mode=oa.NORMAL_F64))
        rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive),
                              hx.HParams(*oa.DEFAULT_PARAMS), 100.0,
                              [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)], n,
                              None, steps, 8, "f64", 37, control_variate="underlying")
        res = pricing._reduce_shards(rq, shard)
        full = c.price_stream_cv(8, n, 37, normal_mode=oa.NORMAL_F64)
        np.testing.assert_allclose(res.sums, full, rtol=1e-12, atol=1e-9)
        want, _ = cv_estimate(full, n, c.offsets, control_means(oa.ASIAN, 100.0, T, steps))
        np.testing.assert_allclose(res.prices, want, rtol=1e-9)
    finally:
        dist.destroy_process_group()


def test_distributed_geometric_sums_with_gloo(hexo_lib):
    """world_size 2 over gloo with the geometric control's sums (5 n_opts doubles per rank)."""
    import torch.multiprocessing as mp
    mp.spawn(_geo_worker, args=(2,), nprocs=2, join=True)


def _geo_worker(rank, world):
    import os
    import sys
    import torch
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    import oracle_api as oa
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import pricing
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29547", rank=rank, world_size=world)
    try:
        T, K, steps, n = [0.5, 1.0], [[100.0], [90.0, 110.0]], 12, 3001
        c = oa.Contract(oa.ASIAN, T, K, steps)

        def shard(rq, begin, count, world_, stats):
            return torch.from_numpy(c.price_stream_geo(
                int(rq.req.seed), int(rq.req.n_paths), int(rq.req.n_streams), begin, count,
                normal_mode=oa.NORMAL_F64))
        chains = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
        rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive),
                              hx.HParams(*oa.DEFAULT_PARAMS), 100.0, chains, n, None, steps, 8,
                              "f64", 29, control_variate="geometric")
        res = pricing._reduce_shards(rq, shard)
        full = c.price_stream_geo(8, n, 29, normal_mode=oa.NORMAL_F64)
        np.testing.assert_allclose(res.sums, full, rtol=1e-12, atol=1e-9)
        plain = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive),
                                 hx.HParams(*oa.DEFAULT_PARAMS), 100.0, chains, n, None, steps, 8,
                                 "f64", 29)
        p0, se0 = pricing._finish(plain, full[:6])
        assert np.all(np.abs(res.prices - p0) < 4 * se0 + 0.02) and np.all(res.stderr < se0)
    finally:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- GPU
CV_CASES = [
    ("asian_chain", oa.ASIAN, [1.0], [[90.0, 100.0, 110.0]], 64, 3000, 96, 0, False),
    ("asian_two_maturities_exact", oa.ASIAN, [0.25, 1.0], [[95.0, 105.0], [100.0]], 20, 1501, 77, 0, True),
    ("euro_three_maturities_philox", oa.EUROPEAN, [0.3, 0.6, 2.0], [[100.0], [100.0], [90.0, 120.0]],
     40, 2000, 70, 1, False),
    ("asian_70_strikes", oa.ASIAN, [0.5], [list(np.linspace(70.0, 130.0, 70))], 16, 900, 64, 0, False),
    ("twelve_maturities", oa.EUROPEAN, [0.1 * (k + 1) for k in range(12)], [[100.0, 101.0]] * 12, 10,
     900, 64, 0, False),
    ("one_path", oa.ASIAN, [1.0], [[100.0]], 16, 1, 1, 0, False),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,payoff,T,K,steps,n_paths,n_streams,rng_mode,exact", CV_CASES,
                         ids=[c[0] for c in CV_CASES])
def test_gpu_cv_sums_match_oracle(gpu, name, payoff, T, K, steps, n_paths, n_streams, rng_mode,
                                  exact):
    import hestonexotics_b200 as hx
    c = oa.Contract(payoff, T, K, steps)
    want = c.price_stream_cv(13, n_paths, n_streams, normal_mode=oa.NORMAL_F64, rng_mode=rng_mode,
                             exact_grid=exact)
    pol = hx.AAsianCallNonAdaptive if payoff == oa.ASIAN else hx.EuropeanCallNonAdaptive
    chains = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
    kw = dict(seed=13, normal_mode="f64", n_streams=n_streams, rng=("shishua", "philox")[rng_mode],
              time_grid="exact" if exact else "reference")
    res = hx.price_full(hx.HQEAnderson(pol), hx.HParams(*oa.DEFAULT_PARAMS), 100.0, chains, n_paths,
                        None, steps, control_variate="underlying", **kw)
    assert res.sums.size == want.size
    # sum c can cancel to near zero: compare against the scale of the summands
    n, nc = c.n_opts, len(T)
    scale = np.concatenate([np.abs(want[:n]), np.abs(want[n:2 * n]), np.abs(want[n:2 * n]),
                            np.sqrt(n_paths * np.abs(want[3 * n + nc:])), np.abs(want[3 * n + nc:])])
    assert np.all(np.abs(res.sums - want) <= 1e-10 * scale + 1e-9)
    # the plain sums are untouched by the control variate
    plain = hx.price_full(hx.HQEAnderson(pol), hx.HParams(*oa.DEFAULT_PARAMS), 100.0, chains,
                          n_paths, None, steps, **kw)
    np.testing.assert_array_equal(res.sums[:2 * n], plain.sums)
    if n_paths > 100:
        wp, wse = cv_estimate(res.sums, n_paths, c.offsets,
                              control_means(payoff, 100.0, T, steps, exact))
        np.testing.assert_allclose(res.prices, wp, rtol=1e-10)
        np.testing.assert_allclose(res.stderr, wse, rtol=1e-8, atol=1e-14)


@pytest.mark.gpu
def test_gpu_cv_very_wide_chains_use_device_accumulators(gpu):
    """6000 strikes over two maturities: 3 n_opts + 2 n_chains accumulators per warp do not fit
    shared memory, the kernel accumulates in device memory.  Same sums as the oracle."""
    import hestonexotics_b200 as hx
    K = list(np.linspace(50, 150, 3000))
    T = [0.25, 0.5]
    c = oa.Contract(oa.EUROPEAN, T, [K, K], 8)
    want = c.price_stream_cv(3, 500, 64, normal_mode=oa.NORMAL_F64)
    r = hx.price_full(hx.HQEAnderson(hx.EuropeanCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                      100.0, [hx.OptionsChain.from_strikes(t, K) for t in T], 500, 6000, 8, seed=3,
                      normal_mode="f64", n_streams=64, control_variate="underlying")
    assert r.sums.size == 3 * 6000 + 4
    assert np.allclose(r.sums, want, rtol=1e-10, atol=1e-7)


@pytest.mark.gpu
def test_gpu_cv_reduces_the_error_and_keeps_the_price(gpu):
    """cfg1-like Asian chain: same paths, standard errors 1.5-5x smaller in and at the money, prices
    within the plain Monte-Carlo error; also through hexo_gpu_price, price_multi and price_batch."""
    import hestonexotics_b200 as hx
    p = hx.HParams(*oa.DEFAULT_PARAMS)
    A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
    Ks = [80.0, 90.0, 100.0, 110.0]
    ch = [hx.OptionsChain.from_strikes(1.0, Ks)]
    plain = hx.price_full(A, p, 100.0, ch, 1_000_000, 4, 252, seed=3)
    cv = hx.price_full(A, p, 100.0, ch, 1_000_000, 4, 252, seed=3, control_variate="underlying")
    ratio = plain.stderr / cv.stderr
    assert ratio[0] > 4 and ratio[1] > 2.5 and ratio[2] > 1.5 and ratio[3] > 1.05, ratio
    assert np.all(np.abs(cv.prices - plain.prices) < 4 * plain.stderr)
    pm, sem = hx.price_multi(A, p, 100.0, ch, 1_000_000, 4, 252, n_gpus=1, seed=3,
                             n_streams=cv.n_streams, control_variate="underlying")
    np.testing.assert_allclose(pm, cv.prices, rtol=1e-12)
    np.testing.assert_allclose(sem, cv.stderr, rtol=1e-9)
    pb, seb, _ = hx.price_batch(A, [p, p], 100.0, ch, 1_000_000, 4, 252, seeds=3,
                                n_streams=cv.n_streams, control_variate="underlying")
    assert np.array_equal(pb[0], cv.prices) and np.array_equal(pb[1], cv.prices)
    assert np.array_equal(seb[0], cv.stderr)


@pytest.mark.gpu
def test_gpu_unknown_control_variate_is_refused(gpu):
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import _lib, pricing
    rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                          100.0, [hx.OptionsChain.from_strikes(1.0, [100.0])], 100, 1, 10, 1,
                          "f32", 32)
    rq.req.control_variate = 7
    sums = np.zeros(8)
    assert gpu.hexo_gpu_price_shard(C.byref(rq.req), 0, 32, sums.ctypes.data_as(_lib.c_double_p),
                                    None) == -1
