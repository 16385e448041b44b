#!/bin/bash
# 2 GPUs: fused exchange on odd shapes (parity vs the unsharded run: max_rel_err_vs_unsharded)
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/sharded_check.py "${@:2}" 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k: d.get(k) for k in ('r_points', 'l_bound', 'steps', 'ms_per_step', 'exchanges_per_step', 'max_rel_err_vs_unsharded', 'norm_err', 'ip_err')})
except Exception as e:
    print('FAILED', e)"; }
run 29571 --r-points 1500 --l-bound 34 --steps 130 --gauge LEN
run 29572 --r-points 7000 --l-bound 26 --steps 70 --gauge LEN
run 29573 --r-points 640 --l-bound 502 --steps 200 --gauge LEN
