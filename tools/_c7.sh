mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_c7.log
python tools/perf_probe.py 1 short 2>&1 | tee gpurun_out/probe_c7.log
python - <<'PY' 2>&1 | tee -a gpurun_out/probe_c7.log
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, hestonexotics_b200 as hx
A=hx.HQEAnderson(hx.AAsianCallNonAdaptive); p=hx.HParams(0.04,0.04,-0.7,2.0,0.5)
for name,T,K,n,steps in (("cfg1-shape asian 252",[1.0],[[100.0]],4_000_000,252),("cfg3 chain 64x8",[0.25*k for k in range(1,9)],[list(np.linspace(70,130,64))]*8,2_000_000,252)):
    ch=[hx.OptionsChain.from_strikes(t,k) for t,k in zip(T,K)]
    best=None
    for i in range(3):
        r=hx.price_full(A,p,100.0,ch,n,None,steps,seed=1)
        if i and (best is None or r.kernel_ms<best.kernel_ms): best=r
    print(f"{name:24s} ms={best.kernel_ms:8.2f} rate={best.path_steps/best.kernel_ms/1e6:7.2f} G/s stepper calls/s={n*best.steps_per_path/best.kernel_ms/1e6:7.2f} G/s")
PY
