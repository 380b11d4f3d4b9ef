mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py > gpurun_out/racecheck_detail.txt 2>&1
grep -E "hazard|Hazard|at |SUMMARY|probe ok" gpurun_out/racecheck_detail.txt | head -60
