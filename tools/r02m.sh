#!/bin/bash
o=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > $o/r02m_bench_n8.json 2> $o/r02m_bench_n8.err; echo "rc=$?"
tail -c 4500 $o/r02m_bench_n8.json; tail -3 $o/r02m_bench_n8.err
