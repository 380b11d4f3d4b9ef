"""Host side of the drop-in: the reference's `HSimulation::price<Scheme>` call
(src/inc/HSimulation.h:61-63, src/HSimulation.tpp:10-51) with the same argument
meaning, served by the CUDA path through the C ABI (include/hexo_gpu.h).

    price(HQEAnderson(AAsianCallNonAdaptive), p, S, all_chains, 100000, n_opts, 1000)

mirrors src/Main.cpp:88.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .types import HParams, OptionsChain, flatten_chains


class AAsianCallNonAdaptive:
    """Arithmetic-average Asian call, trapezoid rule (src/inc/AsianContract.h:14-48)."""
    payoff = _lib.PAYOFF_ASIAN


class EuropeanCallNonAdaptive:
    """European call (src/inc/VanillaContract.h:14-40)."""
    payoff = _lib.PAYOFF_EUROPEAN


@dataclass(frozen=True)
class HQEAnderson:
    """Scheme = Andersen QE stepper + an option policy (src/inc/HSimulation.h:22-47)."""
    policy: type

    @property
    def payoff(self) -> int:
        return self.policy.payoff


_NORMAL_MODES = {"f32": _lib.NORMAL_F32, "as-built": _lib.NORMAL_F32, "f64": _lib.NORMAL_F64,
                 "f32-ppnd7": _lib.NORMAL_F32_PPND7,
                 _lib.NORMAL_F32: _lib.NORMAL_F32, _lib.NORMAL_F64: _lib.NORMAL_F64,
                 _lib.NORMAL_F32_PPND7: _lib.NORMAL_F32_PPND7}
_RNG_MODES = {"shishua": 0, "philox": 1, 0: 0, 1: 1}   # hexo_rng_mode
_GRID_MODES = {"reference": 0, "exact": 1, 0: 0, 1: 1}  # hexo_schedule_mode
_CV_MODES = {None: 0, "none": 0, "underlying": 1, "geometric": 2, 0: 0, 1: 1, 2: 2}  # hexo_control_variate
_DRIFT_MODES = {"reference": 0, "martingale": 1, 0: 0, 1: 1}  # hexo_drift_mode


@dataclass
class PriceResult:
    prices: np.ndarray        # [n_opts], chain-major like the reference's vector
    stderr: np.ndarray        # [n_opts] Monte-Carlo standard errors
    sums: np.ndarray          # [2*n_opts] raw sum(payoff), sum(payoff^2)
    n_paths: int
    n_streams: int
    steps_per_path: int
    path_steps: int
    kernel_ms: float
    grid: int = 0
    block: int = 0


class _Request:
    """Keeps the numpy buffers a hexo_price_request points to alive."""

    def __init__(self, scheme, p: HParams, S: float, all_chains: Sequence[OptionsChain],
                 n_simulations: int, n_opts: Optional[int], steps: int, seed: int, normal_mode,
                 n_streams: int, rng="shishua", time_grid="reference", control_variate=None,
                 drift="reference"):
        if isinstance(scheme, type) and hasattr(scheme, "payoff"):
            scheme = HQEAnderson(scheme)
        self.expiries, self.offsets, self.strikes = flatten_chains(all_chains)
        self.n_opts = int(self.offsets[-1])
        if n_opts is not None and int(n_opts) != self.n_opts:
            # the reference trusts the caller (src/Main.cpp:53-57 computes it); be strict
            raise ValueError(f"n_opts={n_opts} but the chains hold {self.n_opts} options")
        if normal_mode not in _NORMAL_MODES:
            raise ValueError(f"normal_mode must be 'f32', 'f64' or 'f32-ppnd7', got {normal_mode!r}")
        if rng not in _RNG_MODES:
            raise ValueError(f"rng must be 'shishua' or 'philox', got {rng!r}")
        if time_grid not in _GRID_MODES:
            raise ValueError(f"time_grid must be 'reference' or 'exact', got {time_grid!r}")
        if control_variate not in _CV_MODES:
            raise ValueError("control_variate must be None, 'underlying' or 'geometric', got "
                             f"{control_variate!r}")
        if drift not in _DRIFT_MODES:
            raise ValueError(f"drift must be 'reference' or 'martingale', got {drift!r}")
        self.req = _lib.HexoPriceRequest(
            _lib.HexoHParams(*p.as_tuple()), float(S), scheme.payoff, len(self.expiries),
            self.expiries.ctypes.data_as(_lib.c_double_p),
            self.offsets.ctypes.data_as(_lib.c_uint32_p),
            self.strikes.ctypes.data_as(_lib.c_double_p),
            int(n_simulations), int(steps), int(seed), _NORMAL_MODES[normal_mode],
            _RNG_MODES[rng], int(n_streams), _GRID_MODES[time_grid], _CV_MODES[control_variate],
            _DRIFT_MODES[drift])
        cv = self.req.control_variate
        self.n_sums = 5 * self.n_opts if cv == 2 else \
            3 * self.n_opts + 2 * len(self.expiries) if cv else 2 * self.n_opts


def _finish(rq: "_Request", sums: np.ndarray):
    """sums -> (prices, standard errors) with the library's own host routine (hexo_gpu_finish):
    the mean payoff (HSimulation.tpp:40 divides by n_simulations), or the control-variate
    estimate when the request asks for one."""
    lib = _lib.load()
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    assert sums.size == rq.n_sums == lib.hexo_gpu_sums_len(C.byref(rq.req))
    prices, se = np.zeros(rq.n_opts), np.zeros(rq.n_opts)
    _lib.check(lib.hexo_gpu_finish(C.byref(rq.req), sums.ctypes.data_as(_lib.c_double_p),
                                   prices.ctypes.data_as(_lib.c_double_p),
                                   se.ctypes.data_as(_lib.c_double_p)))
    return prices, se


def geometric_asian_means(p: HParams, S: float, all_chains: Sequence[OptionsChain], steps: int,
                          time_grid="reference") -> np.ndarray:
    """E max(G - K, 0) per option: the discretely monitored geometric-Asian call under Heston at
    r = 0 on the time grid of a price<>() call (hexo_heston_geometric_asian; host only)."""
    lib = _lib.load()
    rq = _Request(AAsianCallNonAdaptive, p, S, all_chains, 1, None, steps, 1, "f32", 1,
                  time_grid=time_grid, control_variate="geometric")
    out = np.zeros(rq.n_opts)
    _lib.check(lib.hexo_heston_geometric_asian(C.byref(rq.req), out.ctypes.data_as(_lib.c_double_p)))
    return out


def price_full(scheme, p: HParams, S: float, all_chains: Sequence[OptionsChain],
               n_simulations: int, n_opts: Optional[int], steps: int, *, seed: int = 1,
               normal_mode="f32", n_streams: int = 0, rng="shishua",
               time_grid="reference", control_variate=None, drift="reference",
               device: Optional[int] = None) -> PriceResult:
    """price<Scheme>() on one GPU, returning prices, standard errors and launch statistics."""
    lib = _lib.load()
    if device is not None:
        _lib.check(lib.hexo_gpu_init(int(device)))
    rq = _Request(scheme, p, S, all_chains, n_simulations, n_opts, steps, seed, normal_mode,
                  n_streams, rng, time_grid, control_variate, drift)
    if rq.req.n_streams == 0:
        rq.req.n_streams = lib.hexo_gpu_default_streams(rq.req.n_paths, rq.n_opts, 1)
    sums = np.zeros(rq.n_sums, dtype=np.float64)
    stats = _lib.HexoGpuStats()
    _lib.check(lib.hexo_gpu_price_shard(C.byref(rq.req), 0, rq.req.n_streams,
                                        sums.ctypes.data_as(_lib.c_double_p), C.byref(stats)))
    mean, se = _finish(rq, sums)
    return PriceResult(mean, se, sums, int(n_simulations), int(stats.n_streams),
                       int(stats.steps_per_path), int(stats.path_steps), float(stats.kernel_ms),
                       int(stats.grid), int(stats.block))


def price_multi(scheme, p: HParams, S: float, all_chains: Sequence[OptionsChain],
                n_simulations: int, n_opts: Optional[int], steps: int, *, n_gpus: int = 0,
                seed: int = 1, normal_mode="f32", n_streams: int = 0, rng="shishua",
                time_grid="reference", control_variate=None, drift="reference"):
    """price<Scheme>() spread over several GPUs of THIS process (hexo_gpu_price_multi); returns
    (prices, stderr).  n_gpus = 0 uses every visible device."""
    lib = _lib.load()
    rq = _Request(scheme, p, S, all_chains, n_simulations, n_opts, steps, seed, normal_mode,
                  n_streams, rng, time_grid, control_variate, drift)
    prices, se = np.zeros(rq.n_opts), np.zeros(rq.n_opts)
    _lib.check(lib.hexo_gpu_price_multi(C.byref(rq.req), int(n_gpus),
                                        prices.ctypes.data_as(_lib.c_double_p),
                                        se.ctypes.data_as(_lib.c_double_p), None))
    return prices, se


def price_batch(scheme, params: Sequence[HParams], S: float, all_chains: Sequence[OptionsChain],
                n_simulations: int, n_opts: Optional[int], steps: int, *, seeds=1,
                normal_mode="f32", n_streams: int = 0, rng="shishua", time_grid="reference",
                control_variate=None, drift="reference", n_lanes: int = 0):
    """price<Scheme>() of the same chains for MANY parameter sets in one submission
    (hexo_gpu_price_batch): the shape of Monte-Carlo pricing inside a calibration loop.  `seeds`
    is one seed for all jobs (common random numbers) or one per parameter set.  Returns
    (prices[n_params, n_opts], stderr[n_params, n_opts], batch_ms)."""
    lib = _lib.load()
    params = list(params)
    if not params:
        raise ValueError("price_batch needs at least one parameter set")
    seeds = [int(seeds)] * len(params) if np.isscalar(seeds) else [int(x) for x in seeds]
    if len(seeds) != len(params):
        raise ValueError("one seed per parameter set")
    rqs = [_Request(scheme, p, S, all_chains, n_simulations, n_opts, steps, sd, normal_mode,
                    n_streams, rng, time_grid, control_variate, drift)
           for p, sd in zip(params, seeds)]
    arr = (_lib.HexoPriceRequest * len(rqs))(*[r.req for r in rqs])
    n = rqs[0].n_opts
    prices, se = np.zeros((len(rqs), n)), np.zeros((len(rqs), n))
    stats = (_lib.HexoGpuStats * len(rqs))()
    _lib.check(lib.hexo_gpu_price_batch(arr, len(rqs), int(n_lanes),
                                        prices.ctypes.data_as(_lib.c_double_p),
                                        se.ctypes.data_as(_lib.c_double_p), stats))
    return prices, se, float(stats[0].kernel_ms)


def price(scheme, p: HParams, S: float, all_chains: Sequence[OptionsChain], n_simulations: int,
          n_opts: Optional[int], steps: int, **kw) -> np.ndarray:
    """Drop-in for HSimulation::price<Scheme>: returns the n_opts prices, chain-major."""
    return price_full(scheme, p, S, all_chains, n_simulations, n_opts, steps, **kw).prices


def shard_range(n_streams: int, rank: int, world_size: int):
    """Contiguous stream range of `rank`: the job's streams split as evenly as possible."""
    base, rem = divmod(int(n_streams), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def _gpu_shard(rq: "_Request", begin: int, count: int, world: int, stats):
    """This rank's stream range on its own GPU (torch's current device and stream); returns the
    shard's sums as a device tensor, ready for the all-reduce."""
    import torch
    lib = _lib.load()
    dev = torch.cuda.current_device()
    _lib.check(lib.hexo_gpu_init(dev))
    sums_t = torch.zeros(rq.n_sums, dtype=torch.float64, device=f"cuda:{dev}")
    if count > 0:
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.hexo_gpu_price_shard_device(
            C.byref(rq.req), begin, count, C.c_void_p(sums_t.data_ptr()), C.c_void_p(stream),
            C.byref(stats)))
    return sums_t


def _reduce_shards(rq: "_Request", shard_fn, group=None) -> PriceResult:
    """The multi-rank host logic: split the job's streams over the ranks of `group`, let
    `shard_fn(rq, begin, count, world, stats)` produce this rank's sums (a torch tensor on any
    device), combine them with ONE all-reduce and finish on the host.  price_distributed passes
    the GPU shard; the CPU tests (gloo) pass a stand-in so that this logic runs without a GPU."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rq.req.n_streams == 0:
        rq.req.n_streams = _lib.load().hexo_gpu_default_streams(rq.req.n_paths, rq.n_opts, world)
    stats = _lib.HexoGpuStats()
    begin, count = shard_range(rq.req.n_streams, rank, world)
    sums_t = shard_fn(rq, begin, count, world, stats)
    dist.all_reduce(sums_t, op=dist.ReduceOp.SUM, group=group)
    sums = sums_t.cpu().numpy()
    mean, se = _finish(rq, sums)
    return PriceResult(mean, se, sums, int(rq.req.n_paths), int(rq.req.n_streams),
                       int(stats.steps_per_path), int(rq.req.n_paths) * int(rq.req.steps), 0.0,
                       int(stats.grid), int(stats.block))


def price_distributed(scheme, p: HParams, S: float, all_chains: Sequence[OptionsChain],
                      n_simulations: int, n_opts: Optional[int], steps: int, *, seed: int = 1,
                      normal_mode="f32", n_streams: int = 0, rng="shishua",
                      time_grid="reference", control_variate=None, drift="reference",
                      group=None) -> PriceResult:
    """price<Scheme>() sharded over the ranks of a torch.distributed group.

    Every rank runs a disjoint range of RNG streams on its own GPU and the
    payoff sums (2*n_opts doubles, a few more with a control variate) are combined with ONE
    all-reduce (NCCL over NVLink when the group's backend is nccl).  All ranks return the same
    prices.
    """
    rq = _Request(scheme, p, S, all_chains, n_simulations, n_opts, steps, seed, normal_mode,
                  n_streams, rng, time_grid, control_variate, drift)
    return _reduce_shards(rq, _gpu_shard, group)


def schedule(expiries: Sequence[float], steps: int, time_grid="reference"):
    """Step schedule of a price<>() call (host only): list of (n_steps, h, w, expiry)."""
    lib = _lib.load()
    ex = np.ascontiguousarray(expiries, dtype=np.float64)
    seg = (_lib.HexoSegment * len(ex))()
    fn = lib.hexo_gpu_schedule_exact if _GRID_MODES[time_grid] else lib.hexo_gpu_schedule
    _lib.check(fn(ex.ctypes.data_as(_lib.c_double_p), len(ex), int(steps), seg))
    return [(s.n_steps, s.h, s.w, s.expiry) for s in seg]
