mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_c4.log
python -m pytest tests/test_andersen_cases.py -m gpu -q -s 2>&1 | grep -E "delta|passed|failed" | tee gpurun_out/andersen_test_c4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --paths 20000000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heston_qe -s 1 -c 1 -f -o gpurun_out/prof_r02 \
    python bench.py --paths 4000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_full_r02.log 2>&1
tail -2 gpurun_out/ncu_full_r02.log
