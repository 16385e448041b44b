#!/bin/bash
# 8 GPUs: C5 l-block shards, exchange overlapped with the interior units vs serial; parity vs unsharded in the first run
for env in "X=1" "ION_SERIAL_EXCHANGE=1" "X=1" "ION_SERIAL_EXCHANGE=1"; do
env $env python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/sharded_check.py --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN --no-compare 2>&1 | tail -1 | cut -c1-330
done
env X=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/sharded_check.py --r-points 4096 --l-bound 1024 --steps 64 --gauge VEL 2>&1 | tail -1 | cut -c1-900
env ION_SERIAL_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 tools/sharded_check.py --r-points 4096 --l-bound 1024 --steps 64 --gauge VEL --no-compare 2>&1 | tail -1 | cut -c1-330
