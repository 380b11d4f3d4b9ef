#!/bin/bash
# final measurements of a round on one B200: tests, probe, bench (both arms), launch list, full ncu
# capture of the path kernel, compute-sanitizer.  Outputs land in gpurun_out/.
tag=${1:-final}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/perf_probe.py > gpurun_out/probe_${tag}.log 2>&1; cat gpurun_out/probe_${tag}.log
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 600 gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --paths 20000000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heston_qe -s 1 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py --paths 4000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
for t in memcheck racecheck synccheck initcheck; do
  echo "== $t"; timeout 600 compute-sanitizer --tool $t python tools/sanitize_probe.py 2>&1 | grep -E "COMPUTE-SANITIZER|probe ok|SUMMARY|Error|error" | head -8
done > gpurun_out/sanitizer_${tag}.txt 2>&1
cat gpurun_out/sanitizer_${tag}.txt | tail -16
