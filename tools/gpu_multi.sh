#!/bin/bash
# 2-GPU check: parity tests + timing on one GPU, then the bench contract under torchrun on N GPUs (both arms).
N=${1:-2}
mkdir -p gpurun_out
bash tools/gpu_session.sh
echo "== bench N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-700
tail -3 gpurun_out/bench_n$N.err
echo "== bench reference arm N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>> gpurun_out/bench_n$N.err | tee gpurun_out/bench_ref_n$N.json | cut -c1-300
echo "== bench N=1"
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
