#!/bin/bash
# multi-GPU bench lines: tools/gpu_multi.sh N   (gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
for c in c2 c4; do
  echo "== bench $c N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps 5 --warmup 3 2> gpurun_out/bench_n${N}_$c.err | tee gpurun_out/bench_n${N}_$c.json | cut -c1-400
  tail -2 gpurun_out/bench_n${N}_$c.err
done
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | cut -c1-300
