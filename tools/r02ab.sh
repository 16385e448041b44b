#!/bin/bash
# 2 GPUs: halo exchange fused into the folded length-gauge step (PROG_LEN_STEP_HALO) vs the stand-alone exchange kernel
timeout 300 python -m pytest tests/test_gpu_shards_and_segments.py -m gpu -x -q -k "fused_into" 2>&1 | tail -3
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/sharded_check.py "${@:2}" 2>&1 | tail -1 | cut -c1-900; }
run 29561 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN
ION_FUSED_HALO=0 run 29562 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN --no-compare
run 29563 --r-points 4096 --l-bound 1024 --steps 150 --gauge LEN
ION_FUSED_HALO=0 run 29564 --r-points 4096 --l-bound 1024 --steps 150 --gauge LEN --no-compare
