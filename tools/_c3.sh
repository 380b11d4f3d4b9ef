mkdir -p gpurun_out
python -m pytest tests/test_normals_gpu.py tests/test_config_zscores.py "tests/test_gpu_parity.py::test_fused_kernel_sums_vs_oracle_streams" -m gpu -q -s 2>&1 | grep -E "^\[|K1 normals|fused sums|cfg|passed|failed|Error|assert" | tee gpurun_out/pytest_c3_verbose.log | tail -150
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
