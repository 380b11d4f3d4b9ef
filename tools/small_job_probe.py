"""cfg1-sized jobs (1e5 paths x 252 steps): block size x stream count (development build:
HEXO_BLOCK is only read by libhexo_gpu_dev.so).  usage: HEXO_GPU_LIB=.../libhexo_gpu_dev.so
HEXO_BLOCK=128 python tools/small_job_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib

lib = _lib.load()
_lib.check(lib.hexo_gpu_init(0))
p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
for n in (100_000, 30_000, 200_000, 1_000_000):
    for ns in (0, n, n // 2):
        best = None
        for i in range(5):
            r = hx.price_full(A, p, 100.0, ch, n, 1, 252, seed=1, n_streams=ns)
            if i and (best is None or r.kernel_ms < best.kernel_ms): best = r
        print(f"block={os.environ.get('HEXO_BLOCK','256'):>3s} paths={n:8d} n_streams={best.n_streams:7d} grid={best.grid:5d} "
              f"ms={best.kernel_ms:.3f} rate={best.path_steps/best.kernel_ms/1e6:7.2f} G/s price={best.prices[0]:.4f}", flush=True)
