"""ctypes binding of the C ABI in include/hexo_gpu.h.

The CUDA library is the product: if it is missing this module raises instead of
falling back to anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# HEXO_GPU_LIB: another build of the same CUDA library (A/B comparisons of kernel variants)
LIB_PATH = os.environ.get("HEXO_GPU_LIB") or os.path.join(HERE, "lib", "libhexo_gpu.so")

# names of every function include/hexo_gpu.h declares (checked by tests)
ABI_SYMBOLS = (
    "hexo_gpu_abi_version", "hexo_gpu_init", "hexo_gpu_shutdown", "hexo_gpu_device_count",
    "hexo_gpu_last_error", "hexo_gpu_schedule", "hexo_gpu_price", "hexo_gpu_price_multi",
    "hexo_gpu_price_shard",
    "hexo_gpu_price_shard_device", "hexo_gpu_default_streams", "hexo_gpu_shishua_fill",
    "hexo_gpu_shishua_streams", "hexo_gpu_u64_to_unit", "hexo_gpu_ppnd16", "hexo_gpu_replay",
    "hexo_gpu_measure_fp64_peak", "hexo_gpu_plan_create", "hexo_gpu_plan_launch",
    "hexo_gpu_plan_sums_device", "hexo_gpu_plan_stats", "hexo_gpu_plan_destroy",
    "hexo_gpu_philox4x32", "hexo_gpu_philox_streams", "hexo_gpu_price_batch",
    "hexo_gpu_schedule_exact", "hexo_gpu_sums_len", "hexo_gpu_finish",
    "hexo_gpu_normals_from_words",
)
# host-only semi-analytic benchmark functions of the same library (no hexo_gpu_ prefix)
HOST_SYMBOLS = ("hexo_heston_chf", "hexo_heston_cumulants", "hexo_heston_geometric_asian",
                "hexo_swift_default_params",
                "hexo_swift_price_chain")

HEXO_OK = 0
PAYOFF_ASIAN, PAYOFF_EUROPEAN = 0, 1
NORMAL_F32, NORMAL_F64, NORMAL_F32_PPND7 = 0, 1, 2
RNG_SHISHUA, RNG_PHILOX = 0, 1

c_double_p = C.POINTER(C.c_double)
c_uint32_p = C.POINTER(C.c_uint32)
c_uint64_p = C.POINTER(C.c_uint64)
c_uint8_p = C.POINTER(C.c_uint8)


class HexoHParams(C.Structure):
    _fields_ = [("v_0", C.c_double), ("v_m", C.c_double), ("rho", C.c_double),
                ("kappa", C.c_double), ("sigma", C.c_double)]


class HexoPriceRequest(C.Structure):
    _fields_ = [
        ("p", HexoHParams), ("S", C.c_double), ("payoff", C.c_int32), ("n_chains", C.c_uint32),
        ("expiries", c_double_p), ("strike_offsets", c_uint32_p), ("strikes", c_double_p),
        ("n_paths", C.c_uint64), ("steps", C.c_uint32), ("seed", C.c_uint64),
        ("normal_mode", C.c_int32), ("rng_mode", C.c_int32), ("n_streams", C.c_uint64),
        ("schedule_mode", C.c_int32), ("control_variate", C.c_int32),
        ("drift_mode", C.c_int32),
    ]


class HexoGpuStats(C.Structure):
    _fields_ = [
        ("n_streams", C.c_uint64), ("steps_per_path", C.c_uint64), ("path_steps", C.c_uint64),
        ("grid", C.c_uint32), ("block", C.c_uint32), ("smem_bytes", C.c_uint32),
        ("kernel_launches", C.c_uint32), ("kernel_ms", C.c_float),
    ]


class HexoSegment(C.Structure):
    _fields_ = [("n_steps", C.c_uint32), ("h", C.c_double), ("w", C.c_double),
                ("expiry", C.c_double)]


class HexoSwiftParams(C.Structure):
    _fields_ = [("m", C.c_uint32), ("exp2_m", C.c_uint32), ("sqrt_exp2_m", C.c_double),
                ("lower", C.c_double), ("upper", C.c_double), ("k_1", C.c_int32),
                ("k_2", C.c_int32), ("J", C.c_uint32)]


class HexoGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"hexo_gpu error {code}: {message}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libhexo_gpu.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m hestonexotics_b200.build` "
            "(there is no CPU fallback for the pricing path)")
    lib = C.CDLL(LIB_PATH)
    lib.hexo_gpu_abi_version.restype = C.c_int
    lib.hexo_gpu_init.argtypes = [C.c_int]
    lib.hexo_gpu_last_error.restype = C.c_char_p
    lib.hexo_gpu_schedule.argtypes = [c_double_p, C.c_uint32, C.c_uint32, C.POINTER(HexoSegment)]
    lib.hexo_gpu_schedule_exact.argtypes = lib.hexo_gpu_schedule.argtypes
    lib.hexo_gpu_sums_len.restype = C.c_size_t
    lib.hexo_gpu_sums_len.argtypes = [C.POINTER(HexoPriceRequest)]
    lib.hexo_gpu_finish.argtypes = [C.POINTER(HexoPriceRequest), c_double_p, c_double_p, c_double_p]
    lib.hexo_gpu_price.argtypes = [C.POINTER(HexoPriceRequest), c_double_p, c_double_p,
                                   C.POINTER(HexoGpuStats)]
    lib.hexo_gpu_price_multi.argtypes = [C.POINTER(HexoPriceRequest), C.c_int, c_double_p,
                                         c_double_p, C.POINTER(HexoGpuStats)]
    lib.hexo_gpu_price_shard.argtypes = [C.POINTER(HexoPriceRequest), C.c_uint64, C.c_uint64,
                                         c_double_p, C.POINTER(HexoGpuStats)]
    lib.hexo_gpu_price_shard_device.argtypes = [C.POINTER(HexoPriceRequest), C.c_uint64, C.c_uint64,
                                                C.c_void_p, C.c_void_p, C.POINTER(HexoGpuStats)]
    lib.hexo_gpu_plan_create.argtypes = [C.POINTER(HexoPriceRequest), C.c_uint64, C.c_uint64,
                                         C.POINTER(C.c_void_p)]
    lib.hexo_gpu_plan_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hexo_gpu_plan_sums_device.argtypes = [C.c_void_p]
    lib.hexo_gpu_plan_sums_device.restype = C.c_void_p
    lib.hexo_gpu_plan_stats.argtypes = [C.c_void_p, C.POINTER(HexoGpuStats)]
    lib.hexo_gpu_plan_destroy.argtypes = [C.c_void_p]
    lib.hexo_gpu_default_streams.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
    lib.hexo_gpu_default_streams.restype = C.c_uint64
    lib.hexo_gpu_shishua_fill.argtypes = [c_uint64_p, c_uint8_p, C.c_size_t]
    lib.hexo_gpu_shishua_streams.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, c_uint8_p,
                                             C.c_size_t]
    lib.hexo_gpu_price_batch.argtypes = [C.POINTER(HexoPriceRequest), C.c_uint32, C.c_uint32,
                                         c_double_p, c_double_p, C.POINTER(HexoGpuStats)]
    lib.hexo_gpu_philox4x32.argtypes = [c_uint32_p, c_uint32_p, c_uint32_p, C.c_size_t]
    lib.hexo_gpu_philox_streams.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, c_uint64_p,
                                            C.c_size_t]
    lib.hexo_gpu_u64_to_unit.argtypes = [c_uint64_p, c_double_p, C.c_size_t]
    lib.hexo_gpu_ppnd16.argtypes = [c_double_p, c_double_p, C.c_size_t, C.c_int]
    lib.hexo_gpu_replay.argtypes = [C.POINTER(HexoPriceRequest), c_double_p, C.c_uint64,
                                    C.c_uint32, c_double_p, c_uint32_p]
    lib.hexo_gpu_normals_from_words.argtypes = [c_uint64_p, c_double_p, C.c_size_t, C.c_int]
    lib.hexo_gpu_measure_fp64_peak.argtypes = [c_double_p, C.POINTER(C.c_float)]
    lib.hexo_heston_chf.argtypes = [C.POINTER(HexoHParams), C.c_double, C.c_double, C.c_double,
                                    c_double_p]
    lib.hexo_heston_geometric_asian.argtypes = [C.POINTER(HexoPriceRequest), c_double_p]
    lib.hexo_heston_cumulants.argtypes = [C.POINTER(HexoHParams), C.c_double, c_double_p]
    lib.hexo_swift_default_params.argtypes = [C.POINTER(HexoHParams), C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.POINTER(HexoSwiftParams)]
    lib.hexo_swift_price_chain.argtypes = [C.POINTER(HexoSwiftParams), C.POINTER(HexoHParams),
                                           C.c_double, C.c_double, C.c_double, c_double_p,
                                           C.c_uint32, c_double_p, c_double_p]
    if lib.hexo_gpu_abi_version() != 4:
        raise ImportError("libhexo_gpu.so has an unexpected ABI version; rebuild it")
    _lib = lib
    return lib


def check(code: int) -> int:
    if code < 0:
        raise HexoGpuError(code, load().hexo_gpu_last_error().decode())
    return code
