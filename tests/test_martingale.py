"""Martingale-corrected drift (`drift_mode = HEXO_DRIFT_MARTINGALE`): Andersen's
K0* replaces the constant K0 of the reference's log-spot step (HSimulation.tpp:75,80), so that
E[X' | X, V] = X exactly.  Not the reference's behaviour (SURVEY finding 7: the reference has
no correction).  The oracle restates Andersen (2008) Prop. 9 with the scheme's own a, b^2, p,
beta; an outside anchor -- the martingale identity itself, integrated numerically -- pins the
oracle, and the GPU has to match the oracle like every other mode."""
import ctypes as C

import numpy as np
import pytest

import oracle_api as oa

# (v0, theta, rho, kappa, sigma)
PARAM_SETS = [
    oa.DEFAULT_PARAMS,
    oa.STIFF_PARAMS,
    (0.01, 0.02, -0.3, 0.5, 1.5),   # exponential branch dominates
    (0.09, 0.04, 0.5, 1.0, 0.3),    # positive correlation: A > 0
]


def _k_constants(params, h):
    v0, theta, rho, kappa, eps = params
    K1 = .5 * h * (kappa * rho / eps - .5) - rho / eps
    K2 = .5 * h * (kappa * rho / eps - .5) + rho / eps
    K3 = .5 * h * (1 - rho * rho)
    return K1, K2, K3, K3


def _variance_law(params, h, V):
    """(branch, a, b2) or (branch, p, beta) of the QE step, HSimulation.tpp:58-71."""
    v0, theta, rho, kappa, eps = params
    D = np.exp(-kappa * h)
    m = theta + (V - theta) * D
    s2 = abs(V * eps * eps * D / kappa * (1 - D) + theta * eps * eps / (2 * kappa) * (1 - D) ** 2)
    psi = s2 / m ** 2
    if psi < 1.5:
        b2 = 2 / psi - 1 + np.sqrt(2 / psi * (2 / psi - 1))
        return 0, m / (1 + b2), b2
    return 1, (psi - 1) / (psi + 1), 2 / (m * (psi + 1))


@pytest.mark.parametrize("params", PARAM_SETS)
@pytest.mark.parametrize("h", [1.0 / 1024, 1.0 / 252, 1.0 / 12, 0.5])
def test_k0_star_makes_the_step_a_martingale(params, h):
    """E[exp(ln X' - ln X) | V] = 1: integrate exp(K0* + K1 V + K2 V' + (K3 V + K4 V')/2) over the
    law of V' (Gauss-Hermite in the quadratic branch, quadrature over U in the exponential)."""
    from scipy import integrate
    K1, K2, K3, K4 = _k_constants(params, h)
    z, wz = np.polynomial.hermite_e.hermegauss(160)
    wz = wz / np.sqrt(2 * np.pi)
    seen = set()
    for V in [0.0, 1e-6, 1e-4, 1e-3, 0.01, 0.04, 0.09, 0.3, 1.0]:
        k0, br, ok = oa.k0_star(params, h, V)
        branch, x, y = _variance_law(params, h, V)
        assert br == branch and ok
        seen.add(br)
        if branch == 0:
            a, b2 = x, y
            Vn = a * (np.sqrt(b2) + z) ** 2
            e = float(np.sum(wz * np.exp(k0 + K1 * V + K2 * Vn + .5 * (K3 * V + K4 * Vn))))
        else:
            p, beta = x, y
            f = lambda u: np.exp(k0 + K1 * V + (K2 + .5 * K4) * np.log((1 - p) / (1 - u)) / beta
                                 + .5 * K3 * V)
            tail, _ = integrate.quad(f, p, 1.0, epsabs=1e-13, epsrel=1e-13, limit=400)
            e = p * float(np.exp(k0 + K1 * V + .5 * K3 * V)) + tail
        assert abs(e - 1.0) < 2e-9, (params, h, V, branch, e)
    assert 0 in seen  # every parameter set reaches the quadratic branch somewhere


@pytest.mark.parametrize("params,h,branch", [
    ((3.6, 1.9, 0.92, 8.5, 9.9), 1.1, 1),   # psi >= 1.5 and A >= beta
    ((5.1, 7.4, 0.31, 3.9, 9.2), 3.8, 0),   # psi <  1.5 and 2 A a >= 1
])
def test_k0_star_falls_back_where_the_moment_does_not_exist(params, h, branch):
    """No finite M (only reachable with rho > 0 and year-long steps): the step keeps the
    reference drift K0."""
    v0, theta, rho, kappa, eps = params
    k0, br, ok = oa.k0_star(params, h, v0)
    assert br == branch and not ok
    assert k0 == -rho * kappa * theta / eps * h


def test_oracle_martingale_spot_expectation():
    """Coarse grid (8 steps a year): with the correction the mean terminal spot is S."""
    n, ns = 200_000, 256
    c = oa.Contract(oa.EUROPEAN, [1.0], [[0.0]], 8, drift_mode=1)
    sm, sq = c.price_stream(11, n, ns, normal_mode=oa.NORMAL_F64, exact_grid=True)
    mean = sm[0] / n
    se = np.sqrt((sq[0] / n - mean ** 2) / n)
    assert abs(mean - 100.0) < 4 * se, (mean, se)


def test_request_rejects_unknown_drift(hexo_lib):
    import hestonexotics_b200 as hx
    with pytest.raises(ValueError):
        hx.pricing._Request(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                            100.0, [hx.OptionsChain.from_strikes(1.0, [100.0])], 10, None, 4, 1,
                            "f64", 0, drift="nope")


# ---- GPU ------------------------------------------------------------------------------------

def _chains(T, K):
    import hestonexotics_b200 as hx
    return [hx.OptionsChain.from_strikes(t, k) for t, k in zip(T, K)]


REPLAY = [
    (oa.ASIAN, [1.0], 252, oa.DEFAULT_PARAMS),
    (oa.ASIAN, [0.5, 1.0, 1.5], 40, oa.DEFAULT_PARAMS),
    (oa.EUROPEAN, [0.25, 1.0], 100, oa.STIFF_PARAMS),
    (oa.ASIAN, [1.0], 64, (0.01, 0.02, -0.3, 0.5, 1.5)),      # exponential branch dominates
    (oa.EUROPEAN, [1.0], 12, (0.09, 0.04, 0.5, 1.0, 0.3)),     # A > 0
    (oa.ASIAN, [2.2], 2, (3.6, 1.9, 0.92, 8.5, 9.9)),          # no finite M: reference drift
]


@pytest.mark.gpu
@pytest.mark.parametrize("payoff,expiries,steps,params", REPLAY)
def test_replay_final_values_martingale(gpu, payoff, expiries, steps, params):
    """Same tape, same draws: the stepper with K0* agrees with the oracle to 1e-12 per path."""
    import hestonexotics_b200 as hx
    from hestonexotics_b200 import _lib
    c = oa.Contract(payoff, expiries, [[100.0]] * len(expiries), steps, params, drift_mode=1)
    nsteps = c.steps_to_last_expiry()
    n_paths = 300
    rng = np.random.default_rng(17 + steps)
    tape = np.empty((n_paths, nsteps + 2, 3))
    tape[:, :, 0] = rng.standard_normal((n_paths, nsteps + 2))
    tape[:, :, 1] = rng.random((n_paths, nsteps + 2))
    tape[:, :, 2] = rng.standard_normal((n_paths, nsteps + 2))
    want, used = c.replay(tape)
    plain, _ = oa.Contract(payoff, expiries, [[100.0]] * len(expiries), steps, params).replay(tape)
    if steps > 2:
        assert np.abs(plain - want).max() > 0  # the correction does something
    scheme = hx.HQEAnderson(hx.AAsianCallNonAdaptive if payoff == oa.ASIAN
                            else hx.EuropeanCallNonAdaptive)
    rq = hx.pricing._Request(scheme, hx.HParams(*params), 100.0,
                             _chains(expiries, [[100.0]] * len(expiries)), n_paths, None, steps, 1,
                             "f64", 0, drift="martingale")
    got = np.zeros((n_paths, len(expiries)))
    used = C.c_uint32(0)
    rc = gpu.hexo_gpu_replay(C.byref(rq.req), tape.ctypes.data_as(_lib.c_double_p), n_paths,
                             tape.shape[1], got.ctypes.data_as(_lib.c_double_p), C.byref(used))
    assert rc == 0 and used.value == nsteps, gpu.hexo_gpu_last_error()
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() <= 1e-12, rel.max()


FUSED = [
    # name, payoff, T, K, steps, params, n_paths, n_streams, rng, exact grid
    ("asian_1", oa.ASIAN, [1.0], [[100.0]], 252, oa.DEFAULT_PARAMS, 3000, 96, "shishua", False),
    ("euro_1", oa.EUROPEAN, [1.0], [[90.0, 100.0]], 100, oa.DEFAULT_PARAMS, 3000, 96, "shishua", True),
    ("asian_multi", oa.ASIAN, [0.5, 1.0], [[90.0, 100.0], [100.0, 110.0]], 252, oa.DEFAULT_PARAMS,
     2001, 300, "shishua", False),
    ("asian_12_chains", oa.ASIAN, [0.1 * k for k in range(1, 13)], [[95.0, 105.0]] * 12, 40,
     oa.DEFAULT_PARAMS, 1200, 150, "shishua", True),
    ("exp_branch", oa.ASIAN, [1.0], [[100.0]], 64, (0.01, 0.02, -0.3, 0.5, 1.5), 3000, 128,
     "shishua", False),
    ("stiff", oa.ASIAN, [10.0], [[70.0, 100.0, 130.0]], 2520, oa.STIFF_PARAMS, 200, 64, "shishua",
     False),
    ("philox_euro", oa.EUROPEAN, [0.25, 1.0], [[100.0], [100.0]], 50, oa.DEFAULT_PARAMS, 2000, 77,
     "philox", False),
    ("philox_asian", oa.ASIAN, [1.0], [[100.0]], 128, oa.DEFAULT_PARAMS, 1500, 64, "philox", True),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("f64", 1e-10), ("f32", 1e-4)])
@pytest.mark.parametrize("case", FUSED, ids=[c[0] for c in FUSED])
def test_fused_kernel_martingale_sums_vs_oracle(gpu, case, mode, tol):
    import hestonexotics_b200 as hx
    _, payoff, T, K, steps, params, n_paths, n_streams, rng, exact = case
    c = oa.Contract(payoff, T, K, steps, params, drift_mode=1)
    nm = oa.NORMAL_F64 if mode == "f64" else oa.NORMAL_F32
    sm, sq = c.price_stream(7, n_paths, n_streams, normal_mode=nm, rng_mode=int(rng == "philox"),
                            exact_grid=exact)
    scheme = hx.HQEAnderson(hx.AAsianCallNonAdaptive if payoff == oa.ASIAN
                            else hx.EuropeanCallNonAdaptive)
    res = hx.price_full(scheme, hx.HParams(*params), 100.0, _chains(T, K), n_paths, c.n_opts, steps,
                        seed=7, normal_mode=mode, n_streams=n_streams, rng=rng,
                        time_grid="exact" if exact else "reference", drift="martingale")
    n = c.n_opts
    assert (np.abs(res.sums[:n] - sm) / np.maximum(np.abs(sm), 1e-300)).max() <= tol
    assert (np.abs(res.sums[n:] - sq) / np.maximum(np.abs(sq), 1e-300)).max() <= 2 * tol
    # and it is not the reference drift
    ref = hx.price_full(scheme, hx.HParams(*params), 100.0, _chains(T, K), n_paths, c.n_opts, steps,
                        seed=7, normal_mode=mode, n_streams=n_streams, rng=rng,
                        time_grid="exact" if exact else "reference")
    assert np.abs(ref.sums - res.sums).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("rng", ["shishua", "philox"])
def test_martingale_with_control_variate_matches_oracle(gpu, rng):
    import hestonexotics_b200 as hx
    T, K = [0.5, 1.0], [[90.0, 100.0, 110.0], [100.0]]
    c = oa.Contract(oa.ASIAN, T, K, 64, drift_mode=1)
    want = c.price_stream_cv(3, 2500, 200, normal_mode=oa.NORMAL_F64, rng_mode=int(rng == "philox"),
                             exact_grid=True)
    res = hx.price_full(hx.HQEAnderson(hx.AAsianCallNonAdaptive), hx.HParams(*oa.DEFAULT_PARAMS),
                        100.0, _chains(T, K), 2500, 4, 64, seed=3, normal_mode="f64", n_streams=200,
                        rng=rng, time_grid="exact", control_variate="underlying",
                        drift="martingale")
    assert (np.abs(res.sums - want) / np.maximum(np.abs(want), 1e-300)).max() <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("payoff", ["asian", "european"])
def test_martingale_mean_is_the_spot_on_a_coarse_grid(gpu, payoff):
    """Strike 0 on the exact grid: the payoff is the average (Asian) or the terminal spot
    (European), whose expectation is S when every step is a martingale.  12 steps a year, where
    a drift error of the uncorrected scheme would show; 4e6 paths, SE ~ 0.01."""
    import hestonexotics_b200 as hx
    scheme = hx.HQEAnderson(hx.AAsianCallNonAdaptive if payoff == "asian"
                            else hx.EuropeanCallNonAdaptive)
    p = hx.HParams(0.04, 0.04, -0.9, 0.5, 1.0)  # Andersen's hard case
    res = hx.price_full(scheme, p, 100.0, _chains([1.0], [[0.0]]), 4_000_000, 1, 12, seed=5,
                        normal_mode="f64", time_grid="exact", drift="martingale")
    assert abs(res.prices[0] - 100.0) < 4 * res.stderr[0], (res.prices, res.stderr)
