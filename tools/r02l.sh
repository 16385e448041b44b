#!/bin/bash
# 2 GPUs: l-block shards with the exchange overlapped (one rank per GPU), both gauges, parity vs the unsharded run; then the bench extras
for g in LEN VEL; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/sharded_check.py --r-points 2000 --l-bound 500 --steps 64 --gauge $g 2>&1 | tail -2 | cut -c1-900
ION_SERIAL_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/sharded_check.py --r-points 2000 --l-bound 500 --steps 64 --gauge $g --no-compare 2>&1 | tail -1 | cut -c1-400
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/sharded_check.py --r-points 16384 --l-bound 4096 --steps 40 --gauge LEN --no-compare 2>&1 | tail -1 | cut -c1-400
ION_SERIAL_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/sharded_check.py --r-points 16384 --l-bound 4096 --steps 40 --gauge LEN --no-compare 2>&1 | tail -1 | cut -c1-400
python -m pytest tests -m gpu -x -q -k "shard or peer or config5 or l_block" 2>&1 | tail -3
