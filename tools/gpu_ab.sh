#!/bin/bash
mkdir -p gpurun_out
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_segments.py -m gpu -q -x -s 2>&1 | tail -12
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== segments on one GPU (world 1)"; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_segments_multi.py 10000 8 2>&1 | tail -2
