#!/bin/bash
# 8 GPUs: C5 as odd-cut l-block shards on the folded one-kernel step; parity + efficiency vs the unsharded run on rank 0
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 tools/sharded_check.py "${@:2}" 2>&1 | tail -1 | cut -c1-1000; }
run 29561 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN
ION_SERIAL_EXCHANGE=1 run 29562 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN --no-compare
