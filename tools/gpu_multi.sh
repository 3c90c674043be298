#!/bin/bash
# multi-GPU visit: configs[4] (large-v3, one long clip split over the ranks) + the bench under torchrun
N=${1:-2}; WIN=${2:-30}; TAG=${3:-n$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  tools/run_config5.py --windows $WIN > gpurun_out/config5_$TAG.json 2> gpurun_out/config5_$TAG.err
echo "config5 exit $?"; tail -3 gpurun_out/config5_$TAG.err; cat gpurun_out/config5_$TAG.json
if [ "${WB_SKIP_BENCH:-0}" != "1" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -3 gpurun_out/bench_$TAG.err; python tools/bench_brief.py gpurun_out/bench_$TAG.json
fi
