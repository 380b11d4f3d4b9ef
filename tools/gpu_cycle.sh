#!/bin/bash
# one development cycle on the GPU box: parity tests, throughput probe, ncu capture of the path kernel
tag=${1:-dev}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/quick_gpu.py 2>&1 | tail -11
ncu --set full --clock-control none --import-source on -k regex:heston_qe -s 1 -c 1 -f -o gpurun_out/prof_${tag} python bench.py --paths 4000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
