#!/bin/bash
# final measurements of round 2 on one B200: tests, probes, bench (both arms), launch list, full ncu
# captures of the path kernel (as-built F32 normals and F64 normals), compute-sanitizer.
# Outputs land in gpurun_out/.
tag=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_${tag}.log
grep -q " passed" gpurun_out/pytest_${tag}.log && ! grep -q failed gpurun_out/pytest_${tag}.log || { echo "TESTS FAILED - stopping"; exit 1; }
python tools/perf_probe.py > gpurun_out/probe_${tag}.log 2>&1; cat gpurun_out/probe_${tag}.log
if [ -f hestonexotics_b200/lib/libhexo_gpu_dev.so ]; then
  HEXO_GPU_LIB=$PWD/hestonexotics_b200/lib/libhexo_gpu_dev.so python tools/loop_probe.py 2>&1 | tee gpurun_out/loop_probe_${tag}.log
fi
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 400 gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --paths 20000000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heston_qe -s 1 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py --paths 4000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
ncu --set full --clock-control none -k regex:heston_qe -s 1 -c 1 -f -o gpurun_out/prof_${tag}_f64 \
    python bench.py --paths 4000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras --normal-mode f64 > gpurun_out/ncu_full_f64.log 2>&1
ncu --set full --clock-control none -k regex:heston_qe -s 1 -c 1 -f -o gpurun_out/prof_${tag}_ppnd7 \
    python bench.py --paths 4000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras --normal-mode f32-ppnd7 > gpurun_out/ncu_full_ppnd7.log 2>&1
for t in memcheck racecheck synccheck initcheck; do
  echo "== $t"; timeout 900 compute-sanitizer --tool $t python tools/sanitize_probe.py 2>&1 | grep -E "COMPUTE-SANITIZER|probe ok|SUMMARY|Error|error" | head -8
done > gpurun_out/sanitizer_${tag}.txt 2>&1
cat gpurun_out/sanitizer_${tag}.txt | tail -16
