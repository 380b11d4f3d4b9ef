"""ctypes access to the TEST ORACLE (oracle/libhexo_oracle.so) and, when it has
been built, to the reference's own compiled sources (oracle/_ref/libhexo_ref.so).
Test infrastructure only -- the product never imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libhexo_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libhexo_ref.so")

ASIAN, EUROPEAN = 0, 1
NORMAL_F32, NORMAL_F64 = 0, 1

dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)

DEFAULT_PARAMS = (0.04, 0.04, -0.7, 2.0, 0.5)   # v_0, v_m, rho, kappa, sigma
STIFF_PARAMS = (0.04, 0.04, -0.95, 20.0, 1.0)


class OracleHParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("v_0", "v_m", "rho", "kappa", "sigma")]


class OracleContract(C.Structure):
    _fields_ = [("p", OracleHParams), ("S", C.c_double), ("payoff", C.c_int32),
                ("n_chains", C.c_uint32), ("expiries", dp), ("strike_offsets", u32p),
                ("strikes", dp), ("steps", C.c_uint32), ("drift_mode", C.c_int32)]


def build_oracle():
    src_newer = (not os.path.exists(ORACLE_SO)) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_SO)
        for f in ("hexo_oracle.c", "hexo_oracle.h", "shishua.h", "ppnd16_coef.h"))
    if src_newer:
        subprocess.run(["make", "-C", ORACLE_DIR, "libhexo_oracle.so"], check=True,
                       capture_output=True)


_oracle = None
_ref = None


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        build_oracle()
        o = C.CDLL(ORACLE_SO)
        o.oracle_u64_to_unit.restype = C.c_double
        o.oracle_u64_to_unit.argtypes = [C.c_uint64]
        o.oracle_ppnd16_f64.restype = C.c_double
        o.oracle_ppnd16_f64.argtypes = [C.c_double, C.POINTER(C.c_int)]
        o.oracle_ppnd16_f32.restype = C.c_double
        o.oracle_ppnd16_f32.argtypes = [C.c_double, C.POINTER(C.c_int)]
        o.oracle_normals_from_words.restype = None
        o.oracle_normals_from_words.argtypes = [C.POINTER(C.c_uint64), dp, C.c_size_t, C.c_int]
        o.oracle_rng_new.restype = C.c_void_p
        o.oracle_rng_new.argtypes = [C.c_size_t, C.c_uint, C.c_int]
        o.oracle_rng_grand.restype = C.c_double
        o.oracle_rng_grand.argtypes = [C.c_void_p]
        o.oracle_rng_urand.restype = C.c_double
        o.oracle_rng_urand.argtypes = [C.c_void_p]
        o.oracle_rng_free.argtypes = [C.c_void_p]
        o.oracle_shishua_fill.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_size_t]
        o.oracle_price_ref.argtypes = [C.POINTER(OracleContract), C.c_uint, C.c_uint, C.c_size_t,
                                       C.c_int, dp, dp, dp]
        o.oracle_price_stream.argtypes = [C.POINTER(OracleContract), C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, dp, dp]
        o.oracle_price_stream_rng.argtypes = [C.POINTER(OracleContract), C.c_int, C.c_uint64,
                                              C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                              C.c_int, dp, dp]
        o.oracle_price_stream_exact.argtypes = o.oracle_price_stream_rng.argtypes
        o.oracle_price_stream_cv.argtypes = [C.POINTER(OracleContract), C.c_int, C.c_int,
                                             C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                             C.c_uint64, C.c_int, dp]
        o.oracle_price_stream_geo.argtypes = o.oracle_price_stream_cv.argtypes
        o.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        o.oracle_philox4x32_10.restype = None
        o.oracle_replay.argtypes = [C.POINTER(OracleContract), dp, C.c_uint64, C.c_uint32, dp]
        o.oracle_steps_to_last_expiry.restype = C.c_uint32
        o.oracle_steps_to_last_expiry.argtypes = [C.POINTER(OracleContract)]
        o.oracle_qe_k0_star.restype = C.c_double
        o.oracle_qe_k0_star.argtypes = [C.POINTER(OracleHParams), C.c_double, C.c_double,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _oracle = o
    return _oracle


def k0_star(params, h, V):
    """Andersen's martingale-corrected K0* of one step (oracle_qe_k0_star):
    (K0*, branch 0 quadratic / 1 exponential, corrected flag)."""
    hp = OracleHParams(*[float(x) for x in params])
    br, ok = C.c_int(), C.c_int()
    k0 = oracle().oracle_qe_k0_star(C.byref(hp), float(h), float(V), C.byref(br), C.byref(ok))
    return float(k0), int(br.value), bool(ok.value)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    """The reference's own sources, compiled (oracle/Makefile).  May be absent."""
    global _ref
    if _ref is None:
        r = C.CDLL(REF_SO)
        r.ref_price.argtypes = [C.c_int, dp, C.c_double, C.c_uint, dp, u32p, dp, C.c_uint,
                                C.c_uint, dp]
        r.ref_rng_sequence.argtypes = [C.c_size_t, C.c_uint, C.c_size_t,
                                       C.POINTER(C.c_ubyte), dp]
        _ref = r
    return _ref


class Contract:
    """Owns the numpy buffers an oracle_contract points to."""

    def __init__(self, payoff, expiries, strikes_per_chain, steps, params=DEFAULT_PARAMS, S=100.0,
                 drift_mode=0):
        self.payoff = int(payoff)
        self.params = tuple(float(x) for x in params)
        self.S = float(S)
        self.steps = int(steps)
        self.expiries = np.ascontiguousarray(expiries, dtype=np.float64)
        sizes = [len(s) for s in strikes_per_chain]
        self.offsets = np.zeros(len(sizes) + 1, dtype=np.uint32)
        self.offsets[1:] = np.cumsum(sizes)
        self.strikes = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64)
                                                            for s in strikes_per_chain]))
        self.strikes_per_chain = [list(map(float, s)) for s in strikes_per_chain]
        self.n_opts = int(self.offsets[-1])
        self.c = OracleContract(OracleHParams(*self.params), self.S, self.payoff,
                                len(self.expiries), self.expiries.ctypes.data_as(dp),
                                self.offsets.ctypes.data_as(u32p), self.strikes.ctypes.data_as(dp),
                                self.steps, int(drift_mode))

    # ---- oracle entry points -------------------------------------------------
    def steps_to_last_expiry(self) -> int:
        return int(oracle().oracle_steps_to_last_expiry(C.byref(self.c)))

    def price_ref(self, n_sims, nthreads=1, rand_buf_size=1 << 20, normal_mode=NORMAL_F32):
        pr, sm, sq = (np.zeros(self.n_opts) for _ in range(3))
        rc = oracle().oracle_price_ref(C.byref(self.c), n_sims, nthreads, rand_buf_size,
                                       normal_mode, pr.ctypes.data_as(dp), sm.ctypes.data_as(dp),
                                       sq.ctypes.data_as(dp))
        assert rc == 0, rc
        return pr, sm, sq

    def price_stream(self, seed, n_paths, n_streams, begin=0, count=None,
                     normal_mode=NORMAL_F32, rng_mode=0, exact_grid=False):
        count = n_streams - begin if count is None else count
        sm, sq = np.zeros(self.n_opts), np.zeros(self.n_opts)
        fn = oracle().oracle_price_stream_exact if exact_grid else oracle().oracle_price_stream_rng
        rc = fn(C.byref(self.c), rng_mode, seed, n_paths, n_streams, begin, count, normal_mode,
                sm.ctypes.data_as(dp), sq.ctypes.data_as(dp))
        assert rc == 0, rc
        return sm, sq

    def price_stream_cv(self, seed, n_paths, n_streams, begin=0, count=None,
                        normal_mode=NORMAL_F32, rng_mode=0, exact_grid=False):
        """All sums of a control-variate request: [pf | pf^2 | pf c] per option, [c | c^2] per
        maturity, c = final value - S."""
        count = n_streams - begin if count is None else count
        out = np.zeros(3 * self.n_opts + 2 * len(self.expiries))
        rc = oracle().oracle_price_stream_cv(C.byref(self.c), rng_mode, int(exact_grid), seed,
                                             n_paths, n_streams, begin, count, normal_mode,
                                             out.ctypes.data_as(dp))
        assert rc == 0, rc
        return out

    def price_stream_geo(self, seed, n_paths, n_streams, begin=0, count=None,
                         normal_mode=NORMAL_F32, rng_mode=0, exact_grid=False):
        """All sums of a geometric-control request: [pf | pf^2 | pf c | c | c^2] per option,
        c = max(G - K, 0) with G the geometric average on the arithmetic average's weights."""
        count = n_streams - begin if count is None else count
        out = np.zeros(5 * self.n_opts)
        rc = oracle().oracle_price_stream_geo(C.byref(self.c), rng_mode, int(exact_grid), seed,
                                              n_paths, n_streams, begin, count, normal_mode,
                                              out.ctypes.data_as(dp))
        assert rc == 0, rc
        return out

    def replay(self, tape):
        tape = np.ascontiguousarray(tape, dtype=np.float64)
        n_paths, tape_steps, three = tape.shape
        assert three == 3
        finals = np.zeros((n_paths, len(self.expiries)))
        rc = oracle().oracle_replay(C.byref(self.c), tape.ctypes.data_as(dp), n_paths, tape_steps,
                                    finals.ctypes.data_as(dp))
        assert rc >= 0, rc
        return finals, rc

    # ---- the compiled reference ------------------------------------------------
    def ref_price(self, n_sims, threads=1, normal_mode=NORMAL_F32):
        r = ref()
        r.ref_set_threads(int(threads))
        r.ref_set_normal_mode(int(normal_mode))
        hp = np.asarray(self.params, dtype=np.float64)
        out = np.zeros(self.n_opts)
        rc = r.ref_price(self.payoff, hp.ctypes.data_as(dp), self.S, len(self.expiries),
                         self.expiries.ctypes.data_as(dp), self.offsets.ctypes.data_as(u32p),
                         self.strikes.ctypes.data_as(dp), n_sims, self.steps,
                         out.ctypes.data_as(dp))
        assert rc == 0, rc
        return out


def shishua_bytes(seed4, n_bytes) -> np.ndarray:
    sd = (C.c_uint64 * 4)(*seed4)
    out = np.zeros(n_bytes, dtype=np.uint8)
    rc = oracle().oracle_shishua_fill(sd, out.ctypes.data_as(C.POINTER(C.c_uint8)), n_bytes)
    assert rc == 0
    return out


def u64_to_unit(bits: np.ndarray) -> np.ndarray:
    o = oracle()
    return np.array([o.oracle_u64_to_unit(int(b)) for b in bits], dtype=np.float64)


def ppnd16(p: np.ndarray, normal_mode) -> np.ndarray:
    o = oracle()
    f = o.oracle_ppnd16_f64 if normal_mode == NORMAL_F64 else o.oracle_ppnd16_f32
    ifault = C.c_int(0)
    return np.array([f(float(x), C.byref(ifault)) for x in p], dtype=np.float64)


def normals_from_words(words: np.ndarray, normal_mode) -> np.ndarray:
    """ppnd16(u64_to_unit(word)) for every raw word (RNG.cpp:31,39), vectorised in the oracle."""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    z = np.zeros(len(words))
    oracle().oracle_normals_from_words(words.ctypes.data_as(C.POINTER(C.c_uint64)),
                                       z.ctypes.data_as(dp), len(words), int(normal_mode))
    return z


def rng_sequence(size, seed, kinds, normal_mode=NORMAL_F32) -> np.ndarray:
    """oracle RNG wrapper: kinds[i] != 0 -> get_urand, else get_grand."""
    o = oracle()
    r = o.oracle_rng_new(size, seed, normal_mode)
    assert r
    out = np.array([o.oracle_rng_urand(r) if k else o.oracle_rng_grand(r) for k in kinds])
    o.oracle_rng_free(r)
    return out


def ref_rng_sequence(size, seed, kinds, normal_mode=NORMAL_F32) -> np.ndarray:
    r = ref()
    r.ref_set_normal_mode(int(normal_mode))
    kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
    out = np.zeros(len(kinds))
    rc = r.ref_rng_sequence(size, seed, len(kinds), kinds.ctypes.data_as(C.POINTER(C.c_ubyte)),
                            out.ctypes.data_as(dp))
    assert rc == 0
    return out
