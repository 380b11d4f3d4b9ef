"""CPU tests of the ORACLE: pins the plain-C restatement (oracle/hexo_oracle.c)
against (a) fixtures produced by the reference's own compiled sources
(tests/golden/reference_outputs.json, made by tests/golden/make_golden.py),
(b) that compiled reference live when oracle/_ref exists, and (c) analytic anchors
(AS241 hash sums, scipy's ndtri, closed-form Heston, the reference's rng_test)."""
import json
import os
import re
from decimal import Decimal

import numpy as np
import pytest
from scipy.special import ndtri

import oracle_api as oa
from cases import PRICE_CASES, RNG_CASES
from heston_cf import heston_call

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

with open(os.path.join(HERE, "golden", "reference_outputs.json")) as f:
    GOLDEN = json.load(f)


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs])


# ---- shishua -----------------------------------------------------------------

def test_shishua_round_structure():
    """One round emits 128 bytes; a long fill equals chunked fills of one state."""
    a = oa.shishua_bytes((1, 0, 0, 0), 128 * 64)
    b = oa.shishua_bytes((1, 0, 0, 0), 128 * 8)
    assert np.array_equal(a[:128 * 8], b)
    # different seeds / seed slots give different streams
    seen = {a[:128].tobytes()}
    for sd in [(2, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (1, 1, 0, 0)]:
        s = oa.shishua_bytes(sd, 128).tobytes()
        assert s not in seen
        seen.add(s)


def test_shishua_bits_look_uniform():
    words = oa.shishua_bytes((1, 0, 0, 0), 1 << 20).view(np.uint64)
    bits = np.unpackbits(words.view(np.uint8))
    assert abs(bits.mean() - 0.5) < 2e-3
    u = words.astype(np.float64) / 2.0 ** 64
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3


# ---- uniform map -------------------------------------------------------------

def test_u64_to_unit_edges():
    bits = np.array([0, 1, 2 ** 63, 2 ** 64 - 1, 2 ** 64 - 1024, 2 ** 64 - 1025, 2 ** 53 + 1,
                     0x123456789ABCDEF0], dtype=np.uint64)
    got = oa.u64_to_unit(bits)
    want = bits.astype(np.float64) * 2.0 ** -64   # RN(u64) * 2^-64 (src/RNG.cpp:31)
    assert np.array_equal(got, want)
    assert got[3] == 1.0 and got[0] == 0.0        # [0,1] inclusive


# ---- AS241 -------------------------------------------------------------------

HASH_SUMS = {"AB": Decimal("55.8831928806149014439"), "CD": Decimal("49.33206503301610289036"),
             "EF": Decimal("47.52583317549289671629")}   # as241.f90:45,64,83


def _mantissa_sums(text, pattern):
    sums = {"AB": Decimal(0), "CD": Decimal(0), "EF": Decimal(0)}
    n = 0
    for name, mant in re.findall(pattern, text):
        key = {"A": "AB", "B": "AB", "C": "CD", "D": "CD", "E": "EF", "F": "EF"}[name[0]]
        sums[key] += Decimal(mant)
        n += 1
    return sums, n


def test_ppnd16_hash_sums_oracle_table():
    text = open(os.path.join(ROOT, "oracle", "ppnd16_coef.h")).read()
    sums, n = _mantissa_sums(text, r"#define PPND_([A-F][0-7])\s+([0-9.]+)e[+-]\d+")
    assert n == 45
    assert sums == HASH_SUMS


def test_ppnd16_hash_sums_product_table():
    text = open(os.path.join(ROOT, "hestonexotics_b200", "csrc", "ppnd16.cuh")).read()
    sums, n = _mantissa_sums(text, r"PPND_COEF\(([A-F][0-7]),\s*([0-9.]+)e[+-]\d+\)")
    assert n == 45
    assert sums == HASH_SUMS


def test_central_rational_function_beyond_its_region():
    """The kernel's as-built single-precision mode uses AS241's central rational function up to
    |q| = 0.45 instead of 0.425 (csrc/normals.cuh, F32Split).  What that rests on: with PPND16's
    coefficients the function's own (truncation) error stays far below single precision there --
    <= 3e-10 up to 0.45, i.e. less than 1/400 of half an ulp of a single-precision z in [1, 2) --
    while it is of the order of an ulp at 0.47; and the split constant in the source is 0.45^2 -
    0.180625.  (The GPU test of the transform against the as-built oracle is
    tests/test_normals_gpu.py.)"""
    from scipy.special import ndtri
    text = open(os.path.join(ROOT, "hestonexotics_b200", "csrc", "ppnd16.cuh")).read()
    co = {k: float(v) for k, v in re.findall(r"PPND_COEF\(([AB][0-7]),\s*([0-9.e+-]+)\)", text)}
    assert len(co) == 15

    def central(q):
        r = 0.180625 - q * q
        num = sum(co[f"A{k}"] * r ** k for k in range(8))
        den = 1.0 + sum(co[f"B{k}"] * r ** k for k in range(1, 8))
        return q * num / den

    q = np.linspace(0.0, 0.425, 4001)
    assert np.abs(central(q) - ndtri(0.5 + q)).max() < 2e-15          # AS241's own region
    q = np.linspace(0.425, 0.45, 2001)
    err = np.abs(central(q) - ndtri(0.5 + q))
    assert err.max() < 3e-10 and err.max() < 2.0 ** -24 / 2 / 100     # z in [1.43, 1.65]: ulp 2^-23
    assert abs(central(0.47) - ndtri(0.97)) > 1e-7                    # ... and no further than that
    assert (np.diff(central(np.linspace(0.0, 0.5, 5001))) > 0).all()  # monotone on all of (0, 1/2)
    # the denominator B(r) stays in (0.002, 100) for every |q| <= 1/2: the F64 mode's one
    # reciprocal per pair of draws (normal2_central_f64) relies on it for tail draws' placeholders
    r = 0.180625 - np.linspace(0.0, 0.5, 5001) ** 2
    den = 1.0 + sum(co[f"B{k}"] * r ** k for k in range(1, 8))
    assert den.min() > 2e-3 and den.max() < 100.0
    src = open(os.path.join(ROOT, "hestonexotics_b200", "csrc", "normals.cuh")).read()
    assert "kShift = 0.45f * 0.45f - 0.180625f" in src


def test_ppnd7_hash_sums_product_table():
    """The optional HEXO_NORMAL_F32_PPND7 mode uses AS241's single-precision routine PPND7; its
    coefficients (csrc/normals.cuh) are checked against the hash sums Wichura's paper prints for
    that routine (AB 32.3184577772, CD 15.7614929821), and numerically against the double
    oracle: PPND7 claims "about 1 part in 10**7"."""
    text = open(os.path.join(ROOT, "hestonexotics_b200", "csrc", "normals.cuh")).read()
    body = text[text.index("struct Ppnd7 {"):]
    body = body[:body.index("};")]
    co = {k: Decimal(m) for k, m in re.findall(r"\b([A-D][0-3]) = ([0-9.]+)e[+-]\d+f", body)}
    assert len(co) == 13
    assert sum(v for k, v in co.items() if k[0] in "AB") == Decimal("32.3184577772")
    assert sum(v for k, v in co.items() if k[0] in "CD") == Decimal("15.7614929821")
    val = {k: float(x) for k, x in re.findall(r"\b([A-D][0-3]) = ([0-9.e+-]+)f", body)}
    p = np.concatenate([np.linspace(1e-9, 0.5, 20001), 10.0 ** -np.linspace(2, 10.8, 2000)])
    q = p - 0.5
    r = 0.180625 - q * q
    cen = q * (((val["A3"] * r + val["A2"]) * r + val["A1"]) * r + val["A0"]) / \
        (((val["B3"] * r + val["B2"]) * r + val["B1"]) * r + 1.0)
    rt = np.sqrt(-np.log(p)) - 1.6
    tail = -(((val["C3"] * rt + val["C2"]) * rt + val["C1"]) * rt + val["C0"]) / \
        ((val["D2"] * rt + val["D1"]) * rt + 1.0)
    z = np.where(np.abs(q) <= 0.425, cen, tail)
    ref = oa.ppnd16(p, oa.NORMAL_F64)
    assert (np.abs(z - ref) / np.maximum(1.0, np.abs(ref))).max() < 3e-7


def test_ppnd16_f64_vs_ndtri():
    rng = np.random.default_rng(7)
    p = np.concatenate([rng.random(20000), 10.0 ** -rng.uniform(3, 300, 2000),
                        1 - 10.0 ** -rng.uniform(3, 15, 2000),
                        [0.5, 0.075, 0.925, 0.0749999, 0.9250001, 1e-300]])
    z = oa.ppnd16(p, oa.NORMAL_F64)
    ref = ndtri(p)
    err = np.abs(z - ref) / np.maximum(1.0, np.abs(ref))
    assert err.max() < 5e-15


def test_ppnd16_f32_as_built_accuracy():
    """SURVEY finding 5: the as-built routine differs from double by ~1.4e-6 at most."""
    rng = np.random.default_rng(8)
    p = rng.random(200000)
    d = np.abs(oa.ppnd16(p, oa.NORMAL_F32) - oa.ppnd16(p, oa.NORMAL_F64))
    assert 1e-8 < d.max() < 3e-6
    # every as-built value is a single-precision number
    z = oa.ppnd16(p[:1000], oa.NORMAL_F32)
    assert np.array_equal(z, z.astype(np.float32).astype(np.float64))


def test_ppnd16_ifault_edges():
    for mode in (oa.NORMAL_F32, oa.NORMAL_F64):
        assert oa.ppnd16(np.array([0.0, 1.0]), mode).tolist() == [0.0, 0.0]   # as241.f90:99-103
        assert oa.ppnd16(np.array([0.5]), mode)[0] == 0.0


# ---- RNG wrapper ---------------------------------------------------------------

def test_rng_wrapper_buffer_order():
    """U buffer = first `size` words of the stream, G buffer the next (RNG.cpp:24-26)."""
    size = 128
    words = oa.shishua_bytes((1, 0, 0, 0), 4 * size * 8).view(np.uint64)
    u = oa.rng_sequence(size, 1, [1] * size)
    assert np.array_equal(u, oa.u64_to_unit(words[:size]))
    g = oa.rng_sequence(size, 1, [0] * size, oa.NORMAL_F64)
    assert np.array_equal(g, oa.ppnd16(oa.u64_to_unit(words[size:2 * size]), oa.NORMAL_F64))
    # draining G first makes the third block a G refill
    g2 = oa.rng_sequence(size, 1, [0] * (2 * size), oa.NORMAL_F64)
    assert np.array_equal(g2[size:], oa.ppnd16(oa.u64_to_unit(words[2 * size:3 * size]),
                                               oa.NORMAL_F64))


@pytest.mark.parametrize("name", sorted(RNG_CASES))
@pytest.mark.parametrize("mode", ["f32", "f64"])
def test_rng_wrapper_vs_reference_fixture(name, mode):
    size, seed, pat, reps = RNG_CASES[name]
    kinds = [1 if ch == "u" else 0 for ch in pat] * reps
    nm = oa.NORMAL_F32 if mode == "f32" else oa.NORMAL_F64
    got = oa.rng_sequence(size, seed, kinds, nm)
    want = unhex(GOLDEN["rng"][name][mode])
    assert got.shape == want.shape
    # uniforms are bit-exact; normals of the -ffast-math reference build may differ in the last ulps
    k = np.array(kinds, dtype=bool)
    assert np.array_equal(got[k], want[k])
    assert np.allclose(got[~k], want[~k], rtol=0, atol=1e-6 if mode == "f32" else 1e-13)


def test_rng_moments_like_reference_rng_test():
    """The reference's `hexo -t rng` (src/UnitTest.cpp:525-564) with 2^22 instead of 2^28
    samples; tolerance widened from 1e-4 by sqrt(2^6) accordingly."""
    n = 1 << 22
    o = oa.oracle()
    r = o.oracle_rng_new(1 << 20, 1, oa.NORMAL_F32)
    u = np.empty(n)
    g = np.empty(n)
    for i in range(0, n):
        u[i] = o.oracle_rng_urand(r)
        g[i] = o.oracle_rng_grand(r)
    o.oracle_rng_free(r)
    tol = 1e-4 * 8 * 2
    assert abs(u.mean() - 0.5) < tol and abs(u.var(ddof=1) - 1 / 12) < tol
    assert abs(g.mean()) < tol * 2 and abs(g.var(ddof=1) - 1.0) < tol * 2


# ---- price driver ----------------------------------------------------------------

@pytest.mark.parametrize("name", sorted(PRICE_CASES))
@pytest.mark.parametrize("mode", ["f32", "f64"])
def test_price_vs_reference_fixture(name, mode):
    payoff, T, K, steps, params, n = PRICE_CASES[name]
    c = oa.Contract(payoff, T, K, steps, params)
    nm = oa.NORMAL_F32 if mode == "f32" else oa.NORMAL_F64
    got, _, _ = c.price_ref(n, 1, 1 << 20, nm)
    want = unhex(GOLDEN["prices"][name][mode])
    # same draws, same path logic; the reference build is -ffast-math so not bit-equal
    tol = 1e-9 if mode == "f64" else 5e-5
    assert np.allclose(got, want, rtol=tol, atol=tol * 1e-2)


@pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (needs the reference tree)")
def test_price_vs_live_reference_two_threads():
    """nthreads emulation: seeds 1<<tid and n/nthreads paths per thread (HSimulation.tpp:27-28).
    The live reference's accumulation is racy with >1 thread (:40), so allow for lost updates by
    comparing against the single-thread sums of each emulated thread instead: run the reference
    with 1 thread twice is not possible (seed is fixed), hence only the 1-thread case is exact."""
    c = oa.Contract(oa.ASIAN, [1.0], [[100.0]], 64)
    got, _, _ = c.price_ref(2000, 1)
    want = c.ref_price(2000, threads=1)
    assert np.allclose(got, want, rtol=5e-5)


def test_step_counts_match_survey_table():
    """(T, steps) -> number of stepper calls until the expiry is paid."""
    table = {(1.0, 252): 253, (1.0, 1024): 1024, (1.0, 1000): 1000, (10.0, 2520): 2520,
             (1.0, 365): 366, (1.0, 100): 100}
    for (T, steps), want in table.items():
        assert oa.Contract(oa.ASIAN, [T], [[100.0]], steps).steps_to_last_expiry() == want


def test_european_oracle_vs_closed_form():
    """Stream-convention oracle, f64 normals: MC European within 3.5 SE of closed form (r=0)."""
    c = oa.Contract(oa.EUROPEAN, [1.0], [[100.0]], 252)
    n = 40000
    sm, sq = c.price_stream(seed=1, n_paths=n, n_streams=64, normal_mode=oa.NORMAL_F64)
    mean = sm[0] / n
    se = np.sqrt((sq[0] / n - mean ** 2) / n)
    cf = heston_call(100, 100, 1.0, *oa.DEFAULT_PARAMS, r=0.0)
    assert abs(cf - 7.192552080) < 1e-6
    assert abs(mean - cf) < 3.5 * se


def test_stream_sharding_is_additive():
    c = oa.Contract(oa.ASIAN, [0.5, 1.0], [[95.0, 100.0], [100.0]], 32)
    full = c.price_stream(5, 1003, 17)
    a = c.price_stream(5, 1003, 17, 0, 9)
    b = c.price_stream(5, 1003, 17, 9, 8)
    assert np.allclose(full[0], a[0] + b[0], rtol=1e-13)
    assert np.allclose(full[1], a[1] + b[1], rtol=1e-13)


def test_replay_matches_stream_convention():
    """A tape built from a stream's own words reproduces price_stream's payoffs."""
    c = oa.Contract(oa.ASIAN, [0.25, 0.5], [[100.0], [100.0]], 16)
    nsteps = c.steps_to_last_expiry()
    n_paths = 5
    words = oa.shishua_bytes((9, 0, 0, 0), ((n_paths * nsteps * 2 * 8 + 127) // 128) * 128).view(np.uint64)
    u = oa.u64_to_unit(words[:n_paths * nsteps * 2]).reshape(n_paths, nsteps, 2)
    tape = np.empty((n_paths, nsteps, 3))
    tape[:, :, 0] = oa.ppnd16(u[:, :, 0].ravel(), oa.NORMAL_F64).reshape(n_paths, nsteps)
    tape[:, :, 1] = u[:, :, 0]
    tape[:, :, 2] = oa.ppnd16(u[:, :, 1].ravel(), oa.NORMAL_F64).reshape(n_paths, nsteps)
    finals, used = c.replay(tape)
    assert used == nsteps
    sm, _ = c.price_stream(9, n_paths, 1, normal_mode=oa.NORMAL_F64)
    pay = np.maximum(finals - 100.0, 0.0).sum(axis=0)
    assert np.allclose(pay, sm, rtol=1e-13)


def test_normals_from_words_is_the_scalar_composition():
    """oracle_normals_from_words = ppnd16(u64_to_unit(word)) (RNG.cpp:31,39), both modes."""
    words = oa.shishua_bytes((5, 0, 0, 0), 128 * 16).view(np.uint64)
    words = np.concatenate([words, np.array([0, 2 ** 64 - 1, 2 ** 63], dtype=np.uint64)])
    for mode in (oa.NORMAL_F32, oa.NORMAL_F64):
        z = oa.normals_from_words(words, mode)
        assert np.array_equal(z, oa.ppnd16(oa.u64_to_unit(words), mode))
        assert z[-3] == 0.0 and z[-2] == 0.0 and z[-1] == 0.0


def test_shishua_stream_digest_is_pinned():
    """SHA-256 of the first MiB for the seeds of the reference's first three threads equals the
    committed digest (tests/golden/shishua_sha256.json, written by make_shishua_digest.py, which
    can also diff it against upstream shishua where a network exists)."""
    import hashlib
    import json
    with open(os.path.join(ROOT, "tests", "golden", "shishua_sha256.json")) as f:
        gold = json.load(f)
    for key, want in gold["sha256"].items():
        seed = tuple(int(x) for x in key.split(","))
        got = hashlib.sha256(oa.shishua_bytes(seed, gold["bytes"]).tobytes()).hexdigest()
        assert got == want, key
