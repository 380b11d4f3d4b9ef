"""Development probe: throughput of the FP64 step loop alone (HEXO_NO_REFILL=1) and of the full
kernel at different numbers of warps per SM (block size x blocks).  Needs the development build
of the library (python -m hestonexotics_b200.build --dev; HEXO_GPU_LIB=.../libhexo_gpu_dev.so):
the shipped library ignores these environment variables."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib
lib = _lib.load(); _lib.check(lib.hexo_gpu_init(0))
p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
ch = [hx.OptionsChain.from_strikes(1.0, [100.0])]
for norefill in ("0", "1", "2"):
    os.environ["HEXO_NO_REFILL"] = norefill
    for blk, nblk in ((256, 2), (128, 2)):
        os.environ["HEXO_BLOCK"] = str(blk)
        ns = 148 * nblk * blk
        n = ns * 16
        r = hx.price_full(A, p, 100.0, ch, n, 1, 1024, seed=1, n_streams=ns)
        r = hx.price_full(A, p, 100.0, ch, n, 1, 1024, seed=1, n_streams=ns)
        warps_per_smsp = nblk * blk / 32 / 4
        cyc = r.kernel_ms * 1e-3 * 1.965e9 / (16 * 1024) / warps_per_smsp * 1.0
        print(f"no_refill={norefill} block={blk} blocks/SM={nblk} warps/SMSP={warps_per_smsp:.2f} "
              f"rate={r.path_steps / r.kernel_ms / 1e6:7.1f} G/s  cycles per warp-step per SMSP="
              f"{r.kernel_ms * 1e-3 * 1.965e9 / (16 * 1024) / warps_per_smsp:7.1f}  per-warp latency={r.kernel_ms*1e-3*1.965e9/(16*1024):7.1f}", flush=True)
