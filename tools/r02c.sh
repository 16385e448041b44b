#!/bin/bash
# 2-GPU check of the bench extras (c4 scan split, c5 l-block shards) + the N=1 line
o=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $o/r02c_bench_n2.json 2> $o/r02c_bench_n2.err; echo "rc=$?"; tail -c 3000 $o/r02c_bench_n2.json; tail -5 $o/r02c_bench_n2.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $o/r02c_bench_n1.json 2> $o/r02c_bench_n1.err; echo "rc=$?"; tail -c 3000 $o/r02c_bench_n1.json; tail -5 $o/r02c_bench_n1.err
