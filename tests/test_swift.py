"""Row f1: the host-side SWIFT pricer against the reference's OWN known-answer tests
(tests/golden/swift_kat.json, extracted from src/UnitTest.cpp by tests/golden/make_swift_kat.py),
with the reference's own pass criteria, plus independent cross-checks."""
import json
import os

import numpy as np
import pytest

import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib, swift
from heston_cf import heston_call

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "swift_kat.json")) as f:
    KAT = json.load(f)
P = hx.HParams(*KAT["hparams"])


def kat_params(i):
    m, e, s, lo, up, k1, k2, J = KAT["swift_parameters"][i]
    return _lib.HexoSwiftParams(int(m), int(e), s, lo, up, int(k1), int(k2), int(J))


def test_pricing_kat(hexo_lib):
    """`hexo -t pricing` (src/UnitTest.cpp:219-274): 8 expiries x 5 strikes with the predefined
    swift_parameters; pass criterion sum |diff| < 1e-9."""
    got = []
    for i, (tau, ks) in enumerate(zip(KAT["expiries"], KAT["strikes"])):
        got.extend(swift.swift_price(P, tau, KAT["risk_free"], KAT["S"], ks, kat_params(i)))
    diff = np.abs(np.array(got) - np.array(KAT["prices"])).sum()
    assert diff < 1e-9, diff


def test_gradient_kat(hexo_lib):
    """`hexo -t gradient` (src/UnitTest.cpp:275-497): 200 partials, columns permuted by index_map
    (:478), pass criterion sum |diff| < 1e-9."""
    jac = []
    for i, (tau, ks) in enumerate(zip(KAT["expiries"], KAT["strikes"])):
        _, g = swift.swift_price(P, tau, KAT["risk_free"], KAT["S"], ks, kat_params(i), gradient=True)
        jac.append(g)
    jac = np.concatenate(jac)                       # [40][5] in HParams order
    gold = np.array(KAT["grad"]).reshape(40, 5)
    idx = KAT["grad_index_map"]
    diff = sum(abs(gold[i, idx[j]] - jac[i, j]) for i in range(40) for j in range(5))
    assert diff < 1e-9, diff


def test_default_parameters_reproduce_the_predefined_ones(hexo_lib):
    """SwiftParameters(distr, S, chain) (src/SWIFT.cpp:21-35): the integration bounds match the
    predefined KAT parameter sets for every expiry (to 1e-11: the cumulants come from the
    model's moment cascade here, not from the reference's printed closed forms, so the last two
    digits may differ), and with the truncation precision that yields the KATs' wavelet scale
    m = 5 so do k_1, k_2 and J."""
    for i, (tau, ks) in enumerate(zip(KAT["expiries"], KAT["strikes"])):
        m, e, s, lo, up, k1, k2, J = KAT["swift_parameters"][i]
        q = swift.swift_parameters(P, tau, KAT["risk_free"], KAT["S"], ks[0], ks[-1], 1e-3)
        assert abs(q.lower - lo) < 1e-11 and abs(q.upper - up) < 1e-11
        assert q.exp2_m == 2 ** q.m and abs(q.sqrt_exp2_m - np.sqrt(q.exp2_m)) < 1e-15
        if q.m == int(m):
            assert (q.exp2_m, q.k_1, q.k_2, q.J) == (int(e), int(k1), int(k2), int(J))
        # the release default (1e-7) is finer
        assert swift.swift_parameters(P, tau, KAT["risk_free"], KAT["S"], ks[0], ks[-1]).m >= q.m
    q0 = swift.swift_parameters(P, KAT["expiries"][0], 0.02, 1.0, KAT["strikes"][0][0],
                                KAT["strikes"][0][-1], 1e-3)
    assert (q0.m, q0.k_1, q0.k_2, q0.J) == (5, -58, 50, 128)


def test_gradient_matches_finite_differences(hexo_lib):
    tau, ks = KAT["expiries"][3], KAT["strikes"][3]
    q = kat_params(3)
    _, g = swift.swift_price(P, tau, 0.02, 1.0, ks, q, gradient=True)
    base = np.array(P.as_tuple())
    for j in range(5):
        h = 1e-6
        up, dn = base.copy(), base.copy()
        up[j] += h
        dn[j] -= h
        fd = (swift.swift_price(hx.HParams(*up), tau, 0.02, 1.0, ks, q)
              - swift.swift_price(hx.HParams(*dn), tau, 0.02, 1.0, ks, q)) / (2 * h)
        assert np.allclose(g[:, j], fd, rtol=2e-5, atol=2e-8), j


def test_swift_vs_quadrature_at_r0(hexo_lib):
    """The benchmark the MC European price is compared with (r = 0 because the reference's MC has
    no drift): SWIFT with default parameters agrees with an independent quadrature."""
    p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
    got = swift.swift_price(p, 1.0, 0.0, 100.0, [90.0, 100.0, 110.0])
    want = [heston_call(100, k, 1.0, 0.04, 0.04, -0.7, 2.0, 0.5, 0.0) for k in (90, 100, 110)]
    assert np.allclose(got, want, atol=5e-6)
    assert abs(got[1] - 7.192552080) < 5e-6


def test_chf_basic_properties(hexo_lib):
    v = swift.chf_chf_grad(P, 0.5, 0.0 + 0j)
    assert abs(v[0] - 1.0) < 1e-14                 # chf(0) = 1
    a, b = swift.chf_chf_grad(P, 0.5, 1.3)[0], swift.chf_chf_grad(P, 0.5, -1.3)[0]
    assert abs(a - np.conj(b)) < 1e-14             # Hermitian symmetry
    assert abs(a) < 1.0


def test_bad_arguments(hexo_lib):
    with pytest.raises(_lib.HexoGpuError):
        swift.swift_price(P, -1.0, 0.0, 1.0, [1.0])
    bad = kat_params(0)
    bad.J = 100                                     # not a power of two
    with pytest.raises(_lib.HexoGpuError):
        swift.swift_price(P, 0.5, 0.0, 1.0, [1.0], bad)


def _closed_form_cumulants(p, t):
    """The reference's printed closed forms (src/HDistribution.cpp:90-113), restated here as the
    test oracle for hexo_heston_cumulants (the product derives them from the moment cascade)."""
    v, s2, r, a, k = p
    a2, a3, a4, k2, k3, k4, t2, r2 = a * a, a ** 3, a ** 4, k * k, k ** 3, k ** 4, t * t, r * r
    e = np.exp
    c2 = s2 / (8 * a3) * (-k2 * e(-2 * a * t) + 4 * k * e(-a * t) * (k - 2 * a * r)
                          + 2 * a * t * (4 * a2 + k2 - 4 * a * k * r) + k * (8 * a * r - 3 * k))
    c4 = (3 * k2 * s2) / (64 * a ** 7) * (
        -3 * k4 * e(-4 * a * t)
        - 8 * k2 * e(-3 * a * t) * (2 * a * k * t * (k - 2 * a * r) + 4 * a2 + k2 - 6 * a * k * r)
        - 4 * e(-2 * a * t) * (4 * a2 * k2 * t2 * (k - 2 * a * r) ** 2
                               + 2 * a * k * t * (k3 - 16 * a3 * r - 12 * a * k2 * r + 4 * a2 * k * (3 + 4 * r2))
                               + 8 * a4 - 3 * k4 - 32 * a3 * k * r + 8 * a * k3 * r + 16 * a2 * k2 * r2)
        - 8 * e(-a * t) * (-2 * a2 * k * t2 * (k - 2 * a * r) ** 3
                           - 8 * a * t * (k4 - 7 * a * k3 * r + 4 * a4 * r2 - 8 * a3 * k * r * (1 + r2)
                                          + a2 * k2 * (3 + 14 * r2))
                           - 9 * k4 + 70 * a * k3 * r + 32 * a3 * k * r * (4 + 3 * r2)
                           - 16 * a4 * (1 + 4 * r2) - 4 * a2 * k2 * (9 + 40 * r2))
        + 4 * a * t * (5 * k4 - 40 * a * k3 * r - 32 * a3 * k * r * (3 + 2 * r2) + 16 * a4 * (1 + 4 * r2)
                       + 24 * a2 * k2 * (1 + 4 * r2))
        - 73 * k4 + 544 * a * k3 * r + 128 * a3 * k * r * (7 + 6 * r2) - 32 * a4 * (3 + 16 * r2)
        - 64 * a2 * k2 * (4 + 19 * r2))
    return np.array([-.5 * s2 * t, c2, c4])


def test_cumulants_from_the_moment_cascade_equal_the_closed_forms(hexo_lib):
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(400):
        p = (rng.uniform(0.01, 0.3), rng.uniform(0.01, 0.3), rng.uniform(-0.95, 0.5),
             rng.uniform(0.5, 6.0), rng.uniform(0.1, 1.5))
        tau = rng.uniform(0.05, 5.0)
        got = swift.cumulants(hx.HParams(*p), tau)
        want = _closed_form_cumulants(p, tau)
        worst = max(worst, np.max(np.abs(got - want) / np.abs(want)))
    # both evaluations cancel leading terms for small kappa tau: agreement to ~1e-9 there
    assert worst < 1e-8, worst


def test_chf_is_regular_at_u_zero_and_conjugate_symmetric(hexo_lib):
    """chf(0) = 1 with a zero gradient (the reference's formulas divide by u there), and
    chf(-u) = conj(chf(u)) for real u (what the truncation bound relies on)."""
    v = swift.chf_chf_grad(P, 0.7, 0.0)
    assert abs(v[0] - 1.0) < 1e-15 and np.all(np.abs(v[1:]) < 1e-15)
    for u in (0.3, 7.0, 55.0):
        a, b = swift.chf_chf_grad(P, 0.7, u), swift.chf_chf_grad(P, 0.7, -u)
        assert np.allclose(a, np.conj(b), rtol=1e-13, atol=1e-300)
