#!/bin/bash
# multi-GPU lines: tools/gpu_multi.sh N   (gpurun --gpus N): bench c2 / c4 and one stream cut into N segments
N=${1:-2}
mkdir -p gpurun_out
for c in c2 c4; do
  echo "== bench $c N=$N"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps 5 --warmup 3 2> gpurun_out/bench_n${N}_$c.err | tee gpurun_out/bench_n${N}_$c.json | cut -c1-400
  tail -2 gpurun_out/bench_n${N}_$c.err
done
echo "== one stream in $N segments"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/gpu_segments_multi.py 10000 8 2> gpurun_out/segments_n$N.err | tail -1 | tee gpurun_out/segments_n$N.json
tail -2 gpurun_out/segments_n$N.err
