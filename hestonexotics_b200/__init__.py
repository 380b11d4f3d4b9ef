"""hestonexotics_b200 -- B200-native Heston Monte-Carlo hot path (QE stepper +
shishua + PPND16 + Asian/European payoffs) behind the reference's price<Scheme>()
interface.  The compute path is the CUDA library in lib/libhexo_gpu.so."""
from .types import HParams, Option, OptionsChain, TRADING_DAYS, flatten_chains
from .pricing import (AAsianCallNonAdaptive, EuropeanCallNonAdaptive, HQEAnderson, PriceResult,
                      price, price_full, price_multi, price_batch, price_distributed, schedule, geometric_asian_means,
                      shard_range)

from . import swift  # noqa: E402  (host-side semi-analytic European pricer)

__all__ = [
    "swift",
    "HParams", "Option", "OptionsChain", "TRADING_DAYS", "flatten_chains",
    "AAsianCallNonAdaptive", "EuropeanCallNonAdaptive", "HQEAnderson", "PriceResult",
    "price", "price_full", "price_multi", "price_batch", "price_distributed", "schedule", "shard_range", "geometric_asian_means",
]
