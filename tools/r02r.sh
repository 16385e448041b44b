#!/bin/bash
# 2 GPUs: odd-cut length-gauge shards on the folded one-kernel step (overlapped exchange), parity vs unsharded; VEL regression
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/sharded_check.py "${@:2}" 2>&1 | tail -1 | cut -c1-1000; }
run 29561 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN
ION_SERIAL_EXCHANGE=1 run 29562 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN --no-compare
run 29563 --r-points 2000 --l-bound 500 --steps 64 --gauge LEN
run 29564 --r-points 4096 --l-bound 1024 --steps 64 --gauge VEL
