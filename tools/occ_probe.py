import sys, os
sys.path.insert(0,'/root/repo')
import hestonexotics_b200 as hx
from hestonexotics_b200 import _lib
lib=_lib.load(); _lib.check(lib.hexo_gpu_init(0))
p = hx.HParams(0.04, 0.04, -0.7, 2.0, 0.5)
A = hx.HQEAnderson(hx.AAsianCallNonAdaptive)
ch=[hx.OptionsChain.from_strikes(1.0,[100.0])]
for blk in (256,192,128,64,32):
    os.environ['HEXO_BLOCK']=str(blk)
    for ns in (148*3*blk, 148*2*blk, 148*1*blk):
        r=hx.price_full(A,p,100.0,ch,4_000_000,1,1024,seed=1,n_streams=ns)
        r=hx.price_full(A,p,100.0,ch,4_000_000,1,1024,seed=1,n_streams=ns)
        print(f"block={blk} streams={ns} warps/SM={ns//148//32} ms={r.kernel_ms:.2f} rate={r.path_steps/r.kernel_ms/1e6:.1f} G/s", flush=True)
