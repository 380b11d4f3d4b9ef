"""CPU tests of the drop-in boundary: the C-ABI library builds, loads without a
GPU, exports every function include/hexo_gpu.h declares, fails loudly (no CPU
fallback) and its host-side schedule equals the oracle's literal time loop."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_api as oa
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "hexo_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hexo_gpu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(hexo_lib):
    names = header_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(hexo_lib, n), f"{n} declared in include/hexo_gpu.h but not exported"
    assert sorted(_lib.ABI_SYMBOLS) == names
    text = open(os.path.join(ROOT, "include", "hexo_gpu.h")).read()
    host = sorted(set(re.findall(r"\b(hexo_(?:swift|heston)_\w+)\s*\(", text)))
    assert host == sorted(_lib.HOST_SYMBOLS)
    for n in host:
        assert hasattr(hexo_lib, n), f"{n} declared in include/hexo_gpu.h but not exported"
    assert hexo_lib.hexo_gpu_abi_version() == 4


def test_struct_layout_matches_header(hexo_lib):
    # HParams is five doubles in the reference's order (HDistribution.h:9-24)
    assert [f[0] for f in _lib.HexoHParams._fields_] == ["v_0", "v_m", "rho", "kappa", "sigma"]
    assert C.sizeof(_lib.HexoHParams) == 40
    assert C.sizeof(_lib.HexoSegment) == 32
    assert C.sizeof(_lib.HexoPriceRequest) == 40 + 8 + 8 + 24 + 8 + 8 + 8 + 8 + 8 + 8 + 8


@pytest.mark.parametrize("expiries,steps", [
    ([1.0], 252), ([1.0], 1024), ([1.0], 1000), ([10.0], 2520), ([1.0], 365), ([1.0], 100),
    ([1.0], 1), ([0.5, 1.0], 252), ([0.25, 0.26, 1.0], 100), ([0.25, 0.2501, 0.2502, 0.6], 50),
    ([0.25 * k for k in range(1, 9)], 252), ([1e-3, 5.0], 7), ([0.1, 0.2, 0.3, 0.4, 0.5], 3),
])
def test_schedule_matches_oracle_time_loop(hexo_lib, expiries, steps):
    seg = hx.schedule(expiries, steps)
    total = sum(s[0] for s in seg)
    c = oa.Contract(oa.ASIAN, expiries, [[100.0]] * len(expiries), steps)
    assert total == c.steps_to_last_expiry()
    for (n, h, w, T), Tk in zip(seg, expiries):
        assert T == Tk and h == Tk / steps        # AsianContract.h:35-38
        assert 0.0 <= w <= 1.0 + 1e-9             # interpolation weight inside the last step


def test_schedule_known_values(hexo_lib):
    """SURVEY Appendix B-5."""
    (n, h, w, T), = hx.schedule([1.0], 252)
    assert n == 253 and abs(w - 7.8e-13) < 1e-13
    (n, h, w, T), = hx.schedule([1.0], 1024)
    assert n == 1024 and w == 1.0
    (n, h, w, T), = hx.schedule([10.0], 2520)
    assert n == 2520 and abs(w - 1.0) < 1e-9


def test_schedule_random_against_oracle(hexo_lib):
    rng = np.random.default_rng(3)
    for _ in range(200):
        k = int(rng.integers(1, 6))
        ex = np.cumsum(rng.uniform(1e-3, 1.0, k))
        steps = int(rng.integers(1, 400))
        seg = hx.schedule(ex, steps)
        c = oa.Contract(oa.EUROPEAN, ex, [[100.0]] * k, steps)
        assert sum(s[0] for s in seg) == c.steps_to_last_expiry()


def test_schedule_rejects_bad_input(hexo_lib):
    with pytest.raises(_lib.HexoGpuError) as e:
        hx.schedule([1.0, 0.5], 10)
    assert e.value.code == -2                    # HEXO_ERR_NOT_INCREASING (HSimulation.tpp:15-21)
    with pytest.raises(_lib.HexoGpuError):
        hx.schedule([1.0], 0)
    with pytest.raises(_lib.HexoGpuError):
        hx.schedule([0.0, 1.0], 10)


def test_request_validation_without_gpu(hexo_lib):
    p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
    ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
    with pytest.raises(ValueError):
        hx.price(hx.HQEAnderson(hx.AAsianCallNonAdaptive), p, 100.0, ch, 1000, 2, 252)
    with pytest.raises(ValueError):
        hx.price(hx.HQEAnderson(hx.AAsianCallNonAdaptive), p, 100.0, ch, 1000, 1, 252,
                 normal_mode="f16")


def test_parameter_validation_is_done_before_touching_the_gpu(hexo_lib):
    """Bad model parameters are refused with HEXO_ERR_INVALID_ARGUMENT (the reference would
    divide by kappa / sigma or take log(S) and return NaN prices)."""
    ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
    A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
    good = dict(v_0=0.04, v_m=0.04, rho=-0.7, kappa=2.0, sigma=0.5)
    for bad in (dict(kappa=0.0), dict(sigma=0.0), dict(sigma=-1.0), dict(rho=1.5), dict(v_0=-0.1),
                dict(v_m=float("nan")), dict(kappa=float("inf"))):
        with pytest.raises(_lib.HexoGpuError) as e:
            hx.price(A, hx.HParams(**{**good, **bad}), 100.0, ch, 1000, 1, 16)
        assert e.value.code == -1, bad
    for S in (0.0, -5.0, float("nan")):
        with pytest.raises(_lib.HexoGpuError) as e:
            hx.price(A, hx.HParams(**good), S, ch, 1000, 1, 16)
        assert e.value.code == -1
    with pytest.raises(_lib.HexoGpuError) as e:
        hx.price(A, hx.HParams(**good), 100.0, [hx.OptionsChain.from_strikes(1.0, [float("nan")])],
                 1000, 1, 16)
    assert e.value.code == -1


def test_no_cpu_fallback(hexo_lib):
    """Without a CUDA device the compute entry points must fail, not compute."""
    if hexo_lib.hexo_gpu_device_count() > 0:
        pytest.skip("a GPU is visible here")
    p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
    ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
    with pytest.raises(_lib.HexoGpuError) as e:
        hx.price(hx.HQEAnderson(hx.AAsianCallNonAdaptive), p, 100.0, ch, 1000, 1, 252)
    assert e.value.code == -3                    # HEXO_ERR_NO_DEVICE
    out = np.zeros(128, dtype=np.uint8)
    sd = (C.c_uint64 * 4)(1, 0, 0, 0)
    assert hexo_lib.hexo_gpu_shishua_fill(sd, out.ctypes.data_as(_lib.c_uint8_p), 128) == -3
    assert not out.any()


def test_product_does_not_reference_oracle():
    """The package and the C ABI must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "hestonexotics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libhexo_oracle" not in text and "oracle_api" not in text, f
                assert not re.search(r"#include\s+[\"<].*oracle", text), f


def test_types_mirror_reference():
    ch = hx.OptionsChain.from_strikes(0.5, [90.0, 110.0])
    assert ch.days_to_expiry == int(0.5 * 261) and ch.min_strike == 90.0 and ch.max_strike == 110.0
    ex, off, k = hx.flatten_chains([ch, hx.OptionsChain.from_strikes(1.0, [100.0])])
    assert ex.tolist() == [0.5, 1.0] and off.tolist() == [0, 2, 3] and k.tolist() == [90.0, 110.0, 100.0]
    assert hx.shard_range(10, 0, 4) == (0, 3) and hx.shard_range(10, 3, 4) == (8, 2)
    assert sum(hx.shard_range(75776, r, 8)[1] for r in range(8)) == 75776


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/hexo_gpu.h compiles as strict C99 (no C++ types in the
    signatures) and a C program links against the library and runs its host-only entry points
    (ABI version, step schedule) -- the binding a C / cgo / JNI caller would make."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(r"""
#include <stdio.h>
#include "hexo_gpu.h"
int main(void) {
  double expiries[2] = {0.5, 1.0};
  hexo_segment seg[2];
  if (hexo_gpu_abi_version() != HEXO_GPU_ABI_VERSION) return 1;
  if (hexo_gpu_schedule(expiries, 2, 252, seg) != HEXO_OK) return 2;
  printf("%u %u\n", seg[0].n_steps, seg[1].n_steps);
  return 0;
}
""")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "hestonexotics_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                    "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lhexo_gpu", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    want = hx.schedule([0.5, 1.0], 252)
    assert [int(x) for x in out] == [want[0][0], want[1][0]]
