#!/bin/bash
# multi-GPU measurements of round 2 on one box with N GPUs: the single-process C-ABI path
# (hexo_gpu_price_multi via bench.py --capi-multi N), the one-rank-per-GPU NCCL path (torchrun),
# cfg4 and cfg5, and the multi-device branch of the GPU tests.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m pytest "tests/test_gpu_parity.py::test_single_process_multi_gpu" "tests/test_control_variate.py::test_gpu_cv_reduces_the_error_and_keeps_the_price" -m gpu -q 2>&1 | tail -2
for wl in cfg4 cfg5; do
  python bench.py --capi-multi $N --workload $wl --steps 3 --warmup 1 > gpurun_out/bench_capi_multi_${wl}_n${N}.json 2> gpurun_out/bench_capi_multi_${wl}_n${N}.err
  tail -c 700 gpurun_out/bench_capi_multi_${wl}_n${N}.json; echo
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --workload $wl --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_nccl_${wl}_n${N}.json 2> gpurun_out/bench_nccl_${wl}_n${N}.err
  tail -c 400 gpurun_out/bench_nccl_${wl}_n${N}.json; echo
done
