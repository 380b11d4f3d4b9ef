"""Host-side mirror of the reference's boundary types.

HParams  -> src/inc/HDistribution.h:9-24 (field order v_0, v_m, rho, kappa, sigma)
Option   -> `option`, src/inc/Types.h:26-33
OptionsChain -> `options_chain`, src/inc/Types.h:37-58
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

TRADING_DAYS = float(5 * (365 // 7) + 365 % 7)  # src/inc/BSM.h:9


@dataclass
class HParams:
    v_0: float    # initial variance
    v_m: float    # long-term variance
    rho: float    # correlation between spot and variance
    kappa: float  # mean-reversion rate
    sigma: float  # volatility of variance

    def as_tuple(self):
        return (self.v_0, self.v_m, self.rho, self.kappa, self.sigma)


@dataclass
class Option:
    price: float = 0.0   # ask
    bid: float = 0.0
    strike: float = 0.0
    volume: int = 0


@dataclass
class OptionsChain:
    days_to_expiry: int
    time_to_expiry: float
    options: List[Option] = field(default_factory=list)
    max_strike: float = -np.finfo(np.float64).max
    min_strike: float = np.finfo(np.float64).max

    @classmethod
    def from_strikes(cls, time_to_expiry: float, strikes: Sequence[float]) -> "OptionsChain":
        ch = cls(int(time_to_expiry * TRADING_DAYS), float(time_to_expiry))
        for k in strikes:
            ch.options.append(Option(strike=float(k)))
            ch.max_strike = max(ch.max_strike, float(k))
            ch.min_strike = min(ch.min_strike, float(k))
        return ch


def flatten_chains(all_chains: Sequence[OptionsChain]):
    """std::list<options_chain> -> (expiries, strike_offsets, strikes) arrays."""
    expiries = np.ascontiguousarray([c.time_to_expiry for c in all_chains], dtype=np.float64)
    sizes = [len(c.options) for c in all_chains]
    offsets = np.zeros(len(sizes) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum(sizes)
    strikes = np.ascontiguousarray(
        [o.strike for c in all_chains for o in c.options], dtype=np.float64)
    return expiries, offsets, strikes
