"""Row h: "prices within 3 SE of the reference on every config" (BASELINE.json north_star).

For each of BASELINE's five configurations the reference's OWN code -- HSimulation::price<>
(src/HSimulation.tpp:10-51) with RNG.cpp, compiled verbatim into oracle/_ref/libhexo_ref.so by
oracle/Makefile and called like src/Main.cpp:88 calls it -- prices the contract on ONE host
thread (race-free: the reference's `+=` at HSimulation.tpp:40 is unsynchronised across OpenMP
threads) with >= 1e5 paths, the GPU prices the same contract with >= 100 x as many paths, and
every option must satisfy |z| <= 3 with the combined standard error.

The reference returns no standard error (HSimulation.tpp:39-40 keeps only the mean), so the
reference side's error is the per-path standard deviation the GPU measured for the same option
divided by sqrt(n_ref) -- both sides sample the same payoff distribution.

cfg1 is run at exactly the size BASELINE states (1e5 paths x 252 steps: the reference's own CLI
call).  The GPU and the reference use different random streams (the GPU convention is one
shishua stream per thread, the reference 8 MiB ring buffers per OpenMP thread, SURVEY
Appendix B-6), so this is a statistical test; seeds are fixed, the outcome is deterministic.
"""
import numpy as np
import pytest

import oracle_api as oa
import hestonexotics_b200 as hx

pytestmark = pytest.mark.gpu

ASIAN = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
EURO = hx.HQEAnderson(hx.EuropeanCallNonAdaptive)
K64 = list(np.linspace(70.0, 130.0, 64))

# name, payoff, expiries, strikes per chain, steps, params, reference paths (1 thread), GPU paths
CONFIGS = [
    ("cfg1_asian_100k_x_252", oa.ASIAN, [1.0], [[100.0]], 252, oa.DEFAULT_PARAMS, 100_000, 10_000_000),
    ("cfg2_european_252", oa.EUROPEAN, [1.0], [[100.0]], 252, oa.DEFAULT_PARAMS, 400_000, 40_000_000),
    ("cfg3_chain_64x8_252", oa.ASIAN, [0.25 * k for k in range(1, 9)], [K64] * 8, 252,
     oa.DEFAULT_PARAMS, 100_000, 10_000_000),
    ("cfg4_asian_1024", oa.ASIAN, [1.0], [[100.0]], 1024, oa.DEFAULT_PARAMS, 100_000, 20_000_000),
    ("cfg5_stiff_2520", oa.ASIAN, [10.0], [K64], 2520, oa.STIFF_PARAMS, 100_000, 10_000_000),
]


def zscores(name, payoff, T, K, steps, params, n_ref, n_gpu, seed=1):
    assert oa.have_ref(), "oracle/_ref/libhexo_ref.so is missing: run `make -C oracle` where " \
                          "/root/reference exists (it travels to the GPU box with the snapshot)"
    c = oa.Contract(payoff, T, K, steps, params)
    ref = c.ref_price(n_ref, threads=1)                   # the reference as built (REAL*4 PPND16)
    scheme = ASIAN if payoff == oa.ASIAN else EURO
    chains = [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]
    g = hx.price_full(scheme, hx.HParams(*params), 100.0, chains, n_gpu, c.n_opts, steps, seed=seed)
    sd = g.stderr * np.sqrt(n_gpu)                        # per-path standard deviation
    se = np.hypot(g.stderr, sd / np.sqrt(n_ref))
    live = se > 0                                          # an option nobody ever exercised: 0 == 0
    assert np.array_equal(ref[~live], g.prices[~live])
    z = np.zeros(c.n_opts)
    z[live] = (g.prices[live] - ref[live]) / se[live]
    j = int(np.argmax(np.abs(z)))
    print(f"{name}: {c.n_opts} options, reference {n_ref} paths (1 thread) vs GPU {n_gpu} paths: "
          f"max |z| = {np.abs(z).max():.2f} at option {j} (GPU {g.prices[j]:.5f} +- {g.stderr[j]:.5f}, "
          f"reference {ref[j]:.5f} +- {sd[j] / np.sqrt(n_ref):.5f}); rms z = {np.sqrt(np.mean(z[live] ** 2)):.2f}")
    return z, g, ref


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_config_price_within_3_se_of_the_compiled_reference(gpu, cfg):
    z, g, ref = zscores(*cfg)
    assert np.abs(z).max() <= 3.0, z


def test_cfg4_full_size_price_within_3_se_of_the_reference(gpu):
    """cfg4 at its full 1e9 paths x 1024 steps on this GPU (6 s): the standard error of the GPU
    side is 1.7e-4, so the comparison is limited by the reference sample (2e5 paths, 1 thread,
    20 s); additionally the strike-0 leg reproduces the grid's E[average] = S (1 - 1/steps)."""
    c = oa.Contract(oa.ASIAN, [1.0], [[0.0, 100.0]], 1024)
    n_ref, n = 200_000, 1_000_000_000
    ref = c.ref_price(n_ref, threads=1)
    g = hx.price_full(ASIAN, hx.HParams(*oa.DEFAULT_PARAMS), 100.0,
                      [hx.OptionsChain.from_strikes(1.0, [0.0, 100.0])], n, 2, 1024, seed=1)
    se = np.hypot(g.stderr, g.stderr * np.sqrt(n / n_ref))
    z = (g.prices - ref) / se
    print(f"cfg4 full size: GPU {g.prices} +- {g.stderr}, reference {ref}, z = {z}")
    assert np.abs(z).max() <= 3.0
    assert g.stderr[1] < 2e-4
    # SURVEY finding 6 (+ the QE drift error, < 0.01)
    assert abs(g.prices[0] - 100.0 * (1.0 - 1.0 / 1024)) < 5 * g.stderr[0] + 0.01
