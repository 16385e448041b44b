#!/bin/bash
# 8 GPUs: the bench line with both multi-GPU partitionings (extra.c4_scan, extra.c5_sharded), then C5 with the stand-alone exchange kernel for A/B
o=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > $o/r02ae_bench_n8.json 2> $o/r02ae_bench_n8.err; echo "rc=$?"
tail -c 5000 $o/r02ae_bench_n8.json; tail -3 $o/r02ae_bench_n8.err
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 tools/sharded_check.py "${@:2}" 2>&1 | tail -1 | cut -c1-600; }
ION_FUSED_HALO=0 run 29562 --r-points 16384 --l-bound 4096 --steps 100 --gauge LEN --no-compare
