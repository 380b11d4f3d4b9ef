mkdir -p gpurun_out
python -m pytest tests/test_normals_gpu.py tests/test_config_zscores.py "tests/test_gpu_parity.py::test_fused_kernel_sums_vs_oracle_streams" tests/test_geometric_cv.py -m gpu -q -s 2>&1 | grep -E "^\[|K1 normals|fused sums|cfg|geometric control|passed|failed|Error|assert" | tee gpurun_out/pytest_c3_verbose.log | tail -40
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/perf_probe.py 1 2>&1 | tee gpurun_out/probe_c3.log
python tools/andersen_probe.py 2e7 2>&1 | tee gpurun_out/andersen_c3.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1500 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
