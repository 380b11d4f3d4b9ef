#!/bin/bash
# multi-GPU measurements with BOTH normal modes on one box with N GPUs (round 2, final kernel):
# the one-rank-per-GPU NCCL path (torchrun) and the single-process C-ABI path, cfg4 and cfg5.
# usage: r2_multi_gpu_modes.sh N
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for mode in f64 f32; do
  for wl in cfg4 cfg5; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $N --workload $wl --normal-mode $mode --steps 3 --warmup 3 --no-extras --no-cpu-baseline \
        > gpurun_out/bench_nccl_${wl}_${mode}_n${N}.json 2> gpurun_out/bench_nccl_${wl}_${mode}_n${N}.err
    echo "nccl $wl $mode: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_nccl_${wl}_${mode}_n${N}.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config']['price'], d['clocks'])")"
  done
done
for mode in f64 f32; do
  for wl in cfg4 cfg5; do
    python bench.py --capi-multi $N --workload $wl --normal-mode $mode --steps 3 --warmup 1 \
        > gpurun_out/bench_capi_multi_${wl}_${mode}_n${N}.json 2> gpurun_out/bench_capi_multi_${wl}_${mode}_n${N}.err
    echo "capi $wl $mode: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_capi_multi_${wl}_${mode}_n${N}.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config']['price'], d['same_prices_as_one_device'])")"
  done
done
