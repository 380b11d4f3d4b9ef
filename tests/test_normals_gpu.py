"""Row a9: the inverse-normal transform the fused kernel REALLY runs.

`hexo_gpu_ppnd16` evaluates the scalar AS241 routine (csrc/ppnd16.cuh), which the fused kernel
K1 does not call.  K1 runs `ring_refill` (csrc/path_kernel.cuh): the batched central phase
(two draws per FFMA2, q taken from the high 32 bits of the word, MUFU reciprocal) and the
warp-cooperative tail phase (MUFU lg2 / sqrt / rcp, packed C/D chains) through the
shared-memory ring.  `hexo_gpu_normals_from_words` feeds caller-supplied words through exactly
that code; here it is compared word for word with the oracle's
ppnd16(u64_to_unit(word)) -- RNG::setup_u + RNG::setup_g, src/RNG.cpp:31,39 with
src/as241.f90:85-118.

Stated tolerances
  F64 mode : |dz| <= 5e-15 max(1, |z|)   (AS241 in double on both sides)
  F32 mode : against the as-built single-precision oracle
               |dz| <= 2e-6                         for |z| <= 4   
               |dz| <= 8 x 2^-23 max(1, |z|)        everywhere     (8 single-precision ulps of z:
                                                     beyond |z| = 4.2 one ulp is 4.8e-7)
             and against the DOUBLE oracle the same 8 ulps (the as-built routine itself is up
             to 3.1 ulps = 1.1e-6 away from the true quantile).
             Measured on the B200 (2^24 shishua words): 1.9e-6 for |z| <= 4, 6.1 ulps overall,
             4.7 ulps against the double oracle; the same bounds hold for the optional
             HEXO_NORMAL_F32_PPND7 mode (AS241's single-precision coefficients).
             (The as-built mode evaluates AS241's central rational function up to |q| = 0.45
             instead of 0.425 -- its truncation error there, 2.5e-10, is far below single
             precision; the oracle keeps the reference's 0.425.)
Two single-precision evaluations of the same rational function with different rounding (fused
multiply-add and MUFU approximations on the GPU, separate multiply / add / divide in gfortran's
code) cannot agree better than a few ulps; the measured maximum is printed.
"""
import numpy as np
import pytest

import oracle_api as oa
from hestonexotics_b200 import _lib

pytestmark = pytest.mark.gpu

ULP32 = 2.0 ** -23


def _gpu_normals(gpu, words, mode):
    words = np.ascontiguousarray(words, dtype=np.uint64)
    z = np.zeros(len(words))
    _lib.check(gpu.hexo_gpu_normals_from_words(words.ctypes.data_as(_lib.c_uint64_p),
                                               z.ctypes.data_as(_lib.c_double_p), len(words), mode))
    return z


def _p_to_word(p):
    """the u64 word whose uniform RN(w) 2^-64 is (about) p"""
    return np.uint64(min(int(p * 2.0 ** 64), 2 ** 64 - 1))


def edge_words():
    w = [0, 1, 2, 3, 2 ** 64 - 1, 2 ** 64 - 2, 2 ** 63, 2 ** 63 - 1, 2 ** 63 + 1,
         2 ** 32 - 1, 2 ** 32, 2 ** 32 + 1, 2 ** 53, 2 ** 53 + 1, 2 ** 11, 2 ** 64 - 2 ** 11,
         2 ** 64 - 1024, 2 ** 64 - 1025]
    # |q| = 0.425 (the central / tail split, as241.f90:88) +- a few ulps of the word, both sides;
    # |q| = 0.45 likewise: where the kernel's as-built single-precision mode leaves AS241's central
    # rational function (csrc/normals.cuh, normal2_central_f32)
    for p in (0.075, 0.925, 0.05, 0.95):
        c = int(p * 2.0 ** 64)
        w += [c + d for d in (-2 ** 33, -2 ** 32, -2 ** 12, -1, 0, 1, 2 ** 12, 2 ** 32, 2 ** 33)]
    # r = 5 (the intermediate / far tail split, :105,110): p = exp(-25) = 1.39e-11
    for p in (np.exp(-25.0), 1.0 - np.exp(-25.0)):
        c = min(int(p * 2.0 ** 64), 2 ** 64 - 1)
        w += [min(max(c + d, 0), 2 ** 64 - 1) for d in (-2 ** 22, -2 ** 12, 0, 2 ** 12, 2 ** 22)]
    # far tail p < 1.4e-11 on both sides, down to the last representable words
    for e in range(1, 28):
        w += [2 ** e, 2 ** e + 1, 2 ** 64 - 2 ** e, 3 * 2 ** e]
    # p in the intermediate tail at decades
    for e in range(2, 11):
        w += [int(10.0 ** -e * 2.0 ** 64), 2 ** 64 - 1 - int(10.0 ** -e * 2.0 ** 64)]
    return np.array(w, dtype=np.uint64)


def _check_f32(z, words, label):
    ref32 = oa.normals_from_words(words, oa.NORMAL_F32)
    ref64 = oa.normals_from_words(words, oa.NORMAL_F64)
    assert np.array_equal(z, z.astype(np.float32).astype(np.float64))   # single-precision values
    d = np.abs(z - ref32)
    scale = np.maximum(1.0, np.abs(ref32))
    inner = np.abs(ref32) <= 4.0
    print(f"[{label}] f32: max |dz| = {d.max():.3e}; for |z|<=4: {d[inner].max():.3e}; "
          f"max in ulps of z = {(d / scale).max() / ULP32:.2f}; "
          f"vs double oracle: {(np.abs(z - ref64) / np.maximum(1.0, np.abs(ref64))).max() / ULP32:.2f} ulps "
          f"(as-built oracle vs double: {(np.abs(ref32 - ref64) / np.maximum(1.0, np.abs(ref64))).max() / ULP32:.2f})")
    assert d[inner].max() <= 2e-6
    assert (d / scale).max() <= 8 * ULP32
    assert (np.abs(z - ref64) / np.maximum(1.0, np.abs(ref64))).max() <= 8 * ULP32
    # p in {0, 1}: the reference returns 0 with IFAULT = 1 (as241.f90:99-103)
    return d.max()


@pytest.mark.parametrize("mode", [_lib.NORMAL_F32, _lib.NORMAL_F32_PPND7], ids=["f32", "f32-ppnd7"])
def test_k1_normals_f32_shishua_words(gpu, mode):
    """>= 2^24 words of the reference's own generator (seeds 1<<t like its threads)."""
    worst = 0.0
    for t in range(16):
        words = oa.shishua_bytes((1 << (t % 8), t // 8, 0, 0), 8 << 20).view(np.uint64)
        worst = max(worst, _check_f32(_gpu_normals(gpu, words, mode), words, f"mode {mode} seed {t}"))
    print(f"K1 normals, mode {mode}, 2^24 shishua words: max |dz| vs as-built oracle = {worst:.3e}")


def test_k1_normals_f64_shishua_words(gpu):
    worst = 0.0
    for t in range(16):
        words = oa.shishua_bytes((1 << (t % 8), t // 8, 0, 0), 8 << 20).view(np.uint64)
        z = _gpu_normals(gpu, words, _lib.NORMAL_F64)
        ref = oa.normals_from_words(words, oa.NORMAL_F64)
        worst = max(worst, (np.abs(z - ref) / np.maximum(1.0, np.abs(ref))).max())
    print(f"K1 normals, F64 mode, 2^24 shishua words: max rel |dz| = {worst:.3e}")
    assert worst <= 5e-15


@pytest.mark.parametrize("mode", [_lib.NORMAL_F32, _lib.NORMAL_F64, _lib.NORMAL_F32_PPND7])
def test_k1_normals_edge_words(gpu, mode):
    """Constructed words: 0 and 2^64-1 (p = 0, 1 -> 0), the split points |q| = 0.425 and r = 5
    within a few units of the word, the far tail down to p = 2^-64.  The words are spread over
    many threads and warps (one edge word per ring slot position) so that every slot and both
    planes of the ring see tail draws."""
    edge = edge_words()
    rng = np.random.default_rng(5)
    # place every edge word at several positions of a 32-word chunk, rest: random words
    words = rng.integers(0, 2 ** 64, size=64 * 1024, dtype=np.uint64)
    pos = rng.choice(len(words), size=4 * len(edge), replace=False)
    words[pos] = np.tile(edge, 4)
    z = _gpu_normals(gpu, words, mode)
    if mode != _lib.NORMAL_F64:
        _check_f32(z, words, "edge")
    else:
        ref = oa.normals_from_words(words, oa.NORMAL_F64)
        assert (np.abs(z - ref) / np.maximum(1.0, np.abs(ref))).max() <= 5e-15
    zero = np.isin(words, np.array([0, 2 ** 64 - 1], dtype=np.uint64))
    # RN(2^64 - 1..2^64 - 1024) 2^-64 = 1.0 as well (RNG.cpp:31 rounds to 53 bits)
    one = oa.u64_to_unit(words[pos]) == 1.0
    assert zero.sum() >= 8 and np.all(z[zero] == 0.0) and np.all(z[pos][one] == 0.0)


@pytest.mark.parametrize("mode", [_lib.NORMAL_F32, _lib.NORMAL_F64, _lib.NORMAL_F32_PPND7])
def test_k1_normals_all_tails_overflows_the_list(gpu, mode):
    """A chunk range in which EVERY draw is a tail draw: the warp's tail list (256 entries) cannot
    hold 1024 of them, so the refill falls back to the per-lane loop -- same values."""
    rng = np.random.default_rng(6)
    words = rng.integers(0, int(0.07 * 2 ** 64), size=32 * 1024, dtype=np.uint64)
    words[::2] = np.uint64(2 ** 64 - 1) - words[::2]
    z = _gpu_normals(gpu, words, mode)
    ref = oa.normals_from_words(words, oa.NORMAL_F64 if mode == _lib.NORMAL_F64 else oa.NORMAL_F32)
    tol = 5e-15 if mode == _lib.NORMAL_F64 else 8 * ULP32
    assert (np.abs(z - ref) / np.maximum(1.0, np.abs(ref))).max() <= tol
    assert np.abs(z).min() > 1.4


@pytest.mark.parametrize("n", [1, 7, 31, 32, 33, 8191])
def test_k1_normals_ragged_sizes(gpu, n):
    words = oa.shishua_bytes((3, 1, 0, 0), 128 * 64).view(np.uint64)[:n]
    z = _gpu_normals(gpu, words, _lib.NORMAL_F32)
    ref = oa.normals_from_words(words, oa.NORMAL_F32)
    assert (np.abs(z - ref) / np.maximum(1.0, np.abs(ref))).max() <= 8 * ULP32
