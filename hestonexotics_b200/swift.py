"""Host-side semi-analytic European pricer: the reference's SWIFT
method (src/SWIFT.cpp) and Heston characteristic function (src/HDistribution.cpp),
served by the host code in libhexo_gpu.so.  No GPU involved."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .types import HParams


def chf_chf_grad(p: HParams, tau: float, u: complex) -> np.ndarray:
    """HDistribution::chf_chf_grad(u): [chf, d/dv_0, d/dv_m, d/drho, d/dkappa, d/dsigma]."""
    lib = _lib.load()
    out = np.zeros(12)
    hp = _lib.HexoHParams(*p.as_tuple())
    _lib.check(lib.hexo_heston_chf(C.byref(hp), float(tau), float(np.real(u)), float(np.imag(u)),
                                   out.ctypes.data_as(_lib.c_double_p)))
    return out[0::2] + 1j * out[1::2]


def cumulants(p: HParams, tau: float) -> np.ndarray:
    """Cumulants 1, 2, 4 of the log-return for v_0 := v_m (HDistribution::first/second/
    fourth_order_moment, src/HDistribution.cpp:90-113)."""
    lib = _lib.load()
    out = np.zeros(3)
    hp = _lib.HexoHParams(*p.as_tuple())
    _lib.check(lib.hexo_heston_cumulants(C.byref(hp), float(tau), out.ctypes.data_as(_lib.c_double_p)))
    return out


def swift_parameters(p: HParams, tau: float, risk_free: float, S: float, min_strike: float,
                     max_strike: float, truncation_precision: float = 0.0) -> _lib.HexoSwiftParams:
    """SwiftParameters(distr, S, chain), src/SWIFT.cpp:21-35."""
    lib = _lib.load()
    q = _lib.HexoSwiftParams()
    hp = _lib.HexoHParams(*p.as_tuple())
    _lib.check(lib.hexo_swift_default_params(C.byref(hp), float(tau), float(risk_free), float(S),
                                             float(min_strike), float(max_strike),
                                             float(truncation_precision), C.byref(q)))
    return q


def swift_price(p: HParams, tau: float, risk_free: float, S: float, strikes: Sequence[float],
                params: Optional[_lib.HexoSwiftParams] = None, gradient: bool = False):
    """European call prices of one chain by SWIFT (SWIFT::price_opts); with gradient=True also the
    [n,5] Jacobian in HParams order (SWIFT::price_opts_grad)."""
    lib = _lib.load()
    k = np.ascontiguousarray(strikes, dtype=np.float64)
    if params is None:
        params = swift_parameters(p, tau, risk_free, S, float(k.min()), float(k.max()))
    hp = _lib.HexoHParams(*p.as_tuple())
    prices = np.zeros(len(k))
    grad = np.zeros((len(k), 5)) if gradient else None
    _lib.check(lib.hexo_swift_price_chain(
        C.byref(params), C.byref(hp), float(tau), float(risk_free), float(S),
        k.ctypes.data_as(_lib.c_double_p), len(k), prices.ctypes.data_as(_lib.c_double_p),
        grad.ctypes.data_as(_lib.c_double_p) if gradient else None))
    return (prices, grad) if gradient else prices
