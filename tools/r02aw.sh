#!/bin/bash
# 2 GPUs: the fused exchange with the TMA-staged pair kernel (512-thread CTAs: 2000 x 500) and the multi-GPU tests
timeout 300 python -m pytest tests/test_gpu_shards_and_segments.py -m gpu -x -q -k "fused_into" 2>&1 | tail -3
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/sharded_check.py "${@:2}" 2>&1 | tail -1 | cut -c1-900; }
run 29563 --r-points 2000 --l-bound 500 --steps 150 --gauge LEN
