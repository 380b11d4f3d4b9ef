"""Small pricing calls for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
every kernel variant of the default path on tiny problems."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hestonexotics_b200 as hx

p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
E = hx.HQEAnderson(hx.EuropeanCallNonAdaptive)
one = [hx.OptionsChain.from_strikes(1.0, [100.0])]
many = [hx.OptionsChain.from_strikes(0.1 * (k + 1), list(np.linspace(80, 120, 5))) for k in range(12)]
wide = [hx.OptionsChain.from_strikes(1.0, list(np.linspace(50, 150, 3000)))]
for scheme in (A, E):
    for nm in ("f32", "f64"):
        for rng in ("shishua", "philox"):
            for grid in ("reference", "exact"):
                r = hx.price_full(scheme, p, 100.0, one, 700, None, 20, normal_mode=nm, rng=rng,
                                  time_grid=grid, n_streams=300)
                assert np.isfinite(r.prices).all()
    r = hx.price_full(scheme, p, 100.0, many, 500, None, 6, n_streams=257)   # > 8 segments
    r = hx.price_full(scheme, p, 100.0, wide, 300, None, 4, n_streams=64)    # device accumulators
    r = hx.price_full(scheme, p, 100.0, many, 500, None, 6, n_streams=257, control_variate="underlying")
    r = hx.price_full(scheme, p, 100.0, one, 500, None, 6, n_streams=100, control_variate="underlying")
    r = hx.price_full(scheme, p, 100.0, wide, 300, None, 4, n_streams=64, control_variate="underlying")
    for rng in ("shishua", "philox"):                                         # martingale drift
        r = hx.price_full(scheme, p, 100.0, one, 700, None, 20, n_streams=300, rng=rng,
                          drift="martingale")
        r = hx.price_full(scheme, p, 100.0, many, 500, None, 6, n_streams=257, rng=rng,
                          drift="martingale", control_variate="underlying", time_grid="exact")
        assert np.isfinite(r.prices).all()
# round 2: geometric-Asian control (shared tail list doubles as the warp's buffer of geometric
# averages), the PPND7 normal mode, the word-for-word normal transform kernel
for grid in ("reference", "exact"):
    for drift in ("reference", "martingale"):
        r = hx.price_full(A, p, 100.0, many[:3], 500, None, 6, n_streams=257, drift=drift,
                          time_grid=grid, control_variate="geometric", normal_mode="f64")
        assert np.isfinite(r.prices).all()
r = hx.price_full(A, p, 100.0, wide, 300, None, 4, n_streams=64, control_variate="geometric")
for scheme in (A, E):
    r = hx.price_full(scheme, p, 100.0, many, 500, None, 6, n_streams=257, normal_mode="f32-ppnd7")
    r = hx.price_full(scheme, p, 100.0, one, 700, None, 20, n_streams=300, normal_mode="f32-ppnd7")
from hestonexotics_b200 import _lib
lib = _lib.load()
words = np.random.default_rng(1).integers(0, 2 ** 64, size=5000, dtype=np.uint64)
words[::7] = words[::7] >> np.uint64(5)          # plenty of tail draws
for mode in (0, 1, 2):
    z = np.zeros(len(words))
    _lib.check(lib.hexo_gpu_normals_from_words(words.ctypes.data_as(_lib.c_uint64_p),
                                               z.ctypes.data_as(_lib.c_double_p), len(words), mode))
    assert np.isfinite(z).all()
pr, se, ms = hx.price_batch(A, [p, p, p], 100.0, one, 500, None, 10, n_lanes=2)
print("sanitize probe ok", r.prices[:2], pr[:, 0])
