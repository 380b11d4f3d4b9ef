#!/bin/bash
# round-2 development cycle on one B200: parity tests, then the throughput probe for the product
# library and every variant library under hestonexotics_b200/lib/ (A/B comparisons).
# usage: r2_cycle.sh <tag> [pytest args...]
tag=${1:-dev}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ "$1" != "notest" ]; then
  python -m pytest tests -m gpu -x -q --durations=8 "$@" 2>&1 | tail -25 | tee gpurun_out/pytest_${tag}.log
else shift; fi
for lib in hestonexotics_b200/lib/libhexo_gpu.so hestonexotics_b200/lib/libhexo_gpu_v*.so; do
  [ -f "$lib" ] || continue
  echo "=== $lib"
  HEXO_GPU_LIB=$PWD/$lib python tools/perf_probe.py ${PROBE_SCALE:-1} ${PROBE_SET:-all} 2>&1 | tee -a gpurun_out/probe_${tag}.log
done
if [ -f hestonexotics_b200/lib/libhexo_gpu_dev.so ]; then
  echo "=== step loop alone (development build, HEXO_NO_REFILL)"
  HEXO_GPU_LIB=$PWD/hestonexotics_b200/lib/libhexo_gpu_dev.so python tools/loop_probe.py 2>&1 | tee gpurun_out/loop_probe_${tag}.log
fi
