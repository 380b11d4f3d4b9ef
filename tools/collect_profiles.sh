#!/bin/bash
# Turns the outputs of tools/r2_final_cycle.sh <tag> (gpurun_out/) into the tracked summaries under
# profiles/ (run in the build container: needs ncu to read the .ncu-rep files, no GPU).
# usage: collect_profiles.sh <tag in gpurun_out> <prefix in profiles, e.g. r02>
tag=${1:-r02b}; out=${2:-r02}; G=gpurun_out; P=profiles; T=$(mktemp -d)
WS=128000000   # warp-steps of the profiled launch: 4e6 paths x 1024 steps / 32
cp $G/bench_${tag}.json $P/${out}_bench_1gpu.json
cp $G/bench_ref_${tag}.json $P/${out}_bench_reference_arm.json
cp $G/launches_${tag}.csv $P/${out}_launches_bench.csv
cp $G/probe_${tag}.log $P/${out}_perf_probe.txt
cp $G/loop_probe_${tag}.log $P/${out}_loop_probe_ceilings.txt
cp $G/sanitizer_${tag}.txt $P/${out}_compute_sanitizer.txt
for m in "" _f64 _ppnd7; do
  ncu -i $G/prof_${tag}${m}.ncu-rep --page raw --csv > $T/raw$m.csv 2>/dev/null
  python tools/ncu_keys.py $T/raw$m.csv > $P/${out}_path_kernel_ncu_keys$m.txt
done
cp $T/raw.csv $P/${out}_path_kernel_ncu_raw.csv
ncu -i $G/prof_${tag}.ncu-rep --page source --csv > $T/src.csv 2>/dev/null
python tools/sass_hist.py $T/src.csv $WS > $P/${out}_path_kernel_sass_histogram.txt
python tools/rf_model.py $T/src.csv $WS > $P/${out}_operand_delivery_model.txt
python tools/sass_hot.py $T/src.csv $WS 0.4 > $P/${out}_path_kernel_step_loop_sass.txt
if ncu -i $G/prof_${tag}_f64.ncu-rep --page source --csv > $T/src64.csv 2>/dev/null && [ -s $T/src64.csv ] && grep -q "Instructions Executed" $T/src64.csv; then
  python tools/sass_hist.py $T/src64.csv $WS > $P/${out}_path_kernel_sass_histogram_f64.txt
fi
rm -rf $T
ls -la $P/${out}_*
