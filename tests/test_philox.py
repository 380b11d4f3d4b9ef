"""Optional Philox mode.  The reference has no Philox; the generator is pinned
by the Random123 known-answer vectors (Philox4x32-10), the oracle restates the stream convention
of include/hexo_gpu.h (hexo_rng_mode) and the GPU must match the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key, expected)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
u32p = C.POINTER(C.c_uint32)


def oracle_philox(ctr, key):
    c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    oa.oracle().oracle_philox4x32_10(c, k, o)
    return tuple(o)


def oracle_stream_words(seed, stream, n_steps):
    out = np.empty(2 * n_steps, dtype=np.uint64)
    for n in range(n_steps):
        c = oracle_philox((n & 0xffffffff, n >> 32, stream & 0xffffffff, stream >> 32),
                          (seed & 0xffffffff, seed >> 32))
        out[2 * n] = c[0] | (c[1] << 32)
        out[2 * n + 1] = c[2] | (c[3] << 32)
    return out


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_oracle_philox_known_answers(ctr, key, want):
    assert oracle_philox(ctr, key) == want


def test_oracle_philox_stream_prices_are_sane():
    """Philox-driven oracle prices agree with the shishua-driven ones within Monte-Carlo error."""
    c = oa.Contract(oa.EUROPEAN, [0.5], [[90.0, 100.0, 110.0]], 50)
    n = 40000
    a, a2 = c.price_stream(3, n, 64, normal_mode=oa.NORMAL_F64, rng_mode=1)
    b, b2 = c.price_stream(3, n, 64, normal_mode=oa.NORMAL_F64, rng_mode=0)
    se = np.sqrt((a2 / n - (a / n) ** 2) / n + (b2 / n - (b / n) ** 2) / n)
    assert np.all(np.abs(a / n - b / n) < 4 * se)
    assert not np.array_equal(a, b)


def test_oracle_philox_shards_add_up():
    c = oa.Contract(oa.ASIAN, [0.25, 1.0], [[95.0], [100.0, 105.0]], 20)
    full = c.price_stream(5, 1003, 17, rng_mode=1)
    a = c.price_stream(5, 1003, 17, 0, 9, rng_mode=1)
    b = c.price_stream(5, 1003, 17, 9, 8, rng_mode=1)
    np.testing.assert_allclose(a[0] + b[0], full[0], rtol=1e-13)
    np.testing.assert_allclose(a[1] + b[1], full[1], rtol=1e-13)


# --------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_philox_known_answers(gpu):
    n = len(KAT)
    ctr = np.array([k[0] for k in KAT], dtype=np.uint32).ravel()
    key = np.array([k[1] for k in KAT], dtype=np.uint32).ravel()
    out = np.zeros(4 * n, dtype=np.uint32)
    rc = gpu.hexo_gpu_philox4x32(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p),
                                 out.ctypes.data_as(u32p), n)
    assert rc == 0
    assert [tuple(int(x) for x in out[4 * i:4 * i + 4]) for i in range(n)] == [k[2] for k in KAT]


@pytest.mark.gpu
def test_gpu_philox_blocks_match_oracle(gpu):
    rng = np.random.default_rng(5)
    n = 4096
    ctr = rng.integers(0, 2 ** 32, size=4 * n, dtype=np.uint32)
    key = rng.integers(0, 2 ** 32, size=2 * n, dtype=np.uint32)
    out = np.zeros(4 * n, dtype=np.uint32)
    assert gpu.hexo_gpu_philox4x32(ctr.ctypes.data_as(u32p), key.ctypes.data_as(u32p),
                                   out.ctypes.data_as(u32p), n) == 0
    for i in range(0, n, 37):
        want = oracle_philox(tuple(int(x) for x in ctr[4 * i:4 * i + 4]),
                             tuple(int(x) for x in key[2 * i:2 * i + 2]))
        assert tuple(int(x) for x in out[4 * i:4 * i + 4]) == want


@pytest.mark.gpu
def test_gpu_philox_stream_words_match_oracle(gpu):
    seed, first, n_streams, words = (0x1234567 << 32) | 99, (1 << 32) + 5, 3, 64
    out = np.zeros(n_streams * words, dtype=np.uint64)
    assert gpu.hexo_gpu_philox_streams(seed, first, n_streams,
                                       out.ctypes.data_as(C.POINTER(C.c_uint64)), words) == 0
    for s in range(n_streams):
        np.testing.assert_array_equal(out[s * words:(s + 1) * words],
                                      oracle_stream_words(seed, first + s, words // 2))


PHILOX_CASES = [
    ("asian_atm", oa.ASIAN, [1.0], [[100.0]], 64, 3000, 96, "f64"),
    ("asian_f32", oa.ASIAN, [1.0], [[100.0]], 64, 3000, 96, "f32"),
    ("euro_chain", oa.EUROPEAN, [0.5], [[80.0, 100.0, 120.0]], 50, 2500, 130, "f64"),
    ("asian_two_maturities", oa.ASIAN, [0.25, 1.0], [[95.0, 105.0], [100.0]], 20, 1501, 77, "f64"),
    ("feller_violated", oa.ASIAN, [1.0], [[100.0]], 40, 2000, 64, "f64"),
    ("twelve_chains", oa.EUROPEAN, [0.1 * (k + 1) for k in range(12)],
     [[100.0]] * 12, 10, 900, 64, "f64"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,payoff,T,K,steps,n_paths,n_streams,nm", PHILOX_CASES,
                         ids=[c[0] for c in PHILOX_CASES])
def test_gpu_philox_fused_sums_match_oracle(gpu, name, payoff, T, K, steps, n_paths, n_streams, nm):
    import hestonexotics_b200 as hx
    params = oa.DEFAULT_PARAMS
    if name == "feller_violated":   # frequent psi >= 1.5: exercises the uniform of the same word
        params = (0.04, 0.04, -0.7, 0.5, 1.0)
    c = oa.Contract(payoff, T, K, steps, params)
    omode = oa.NORMAL_F64 if nm == "f64" else oa.NORMAL_F32
    sm, sq = c.price_stream(11, n_paths, n_streams, normal_mode=omode, rng_mode=1)
    p = hx.HParams(*params)
    pol = hx.AAsianCallNonAdaptive if payoff == oa.ASIAN else hx.EuropeanCallNonAdaptive
    chains = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
    res = hx.price_full(hx.HQEAnderson(pol), p, c.S, chains, n_paths, None, steps, seed=11,
                        normal_mode=nm, n_streams=n_streams, rng="philox")
    tol = 1e-10 if nm == "f64" else 2e-4   # f32: the single-precision transform differs in ulps
    np.testing.assert_allclose(res.sums[:c.n_opts], sm, rtol=tol, atol=tol)
    np.testing.assert_allclose(res.sums[c.n_opts:], sq, rtol=10 * tol, atol=tol)


@pytest.mark.gpu
def test_gpu_philox_price_within_error_of_closed_form(gpu):
    import hestonexotics_b200 as hx
    from heston_cf import heston_call
    p = hx.HParams(*oa.DEFAULT_PARAMS)
    Ks = [90.0, 100.0, 110.0]
    res = hx.price_full(hx.HQEAnderson(hx.EuropeanCallNonAdaptive), p, 100.0,
                        [hx.OptionsChain.from_strikes(1.0, Ks)], 4_000_000, 3, 128, seed=3,
                        rng="philox")
    want = np.array([heston_call(100.0, k, 1.0, *oa.DEFAULT_PARAMS) for k in Ks])
    assert np.all(np.abs(res.prices - want) < 4 * res.stderr + 0.02)   # + QE discretisation bias
    shi = hx.price_full(hx.HQEAnderson(hx.EuropeanCallNonAdaptive), p, 100.0,
                        [hx.OptionsChain.from_strikes(1.0, Ks)], 4_000_000, 3, 128, seed=3)
    assert np.all(np.abs(res.prices - shi.prices) < 5 * np.hypot(res.stderr, shi.stderr))
    assert not np.array_equal(res.prices, shi.prices)


@pytest.mark.gpu
def test_gpu_unknown_rng_mode_is_refused(gpu):
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import _lib, pricing
    rq = pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                          100.0, [hx.OptionsChain.from_strikes(1.0, [100.0])], 100, 1, 10, 1,
                          "f32", 32)
    rq.req.rng_mode = 7
    sums = np.zeros(2)
    rc = gpu.hexo_gpu_price_shard(C.byref(rq.req), 0, 32, sums.ctypes.data_as(_lib.c_double_p), None)
    assert rc == -1
