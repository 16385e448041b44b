#!/bin/bash
# the driver's round-end sequence on 4 GPUs: bench at N = 4 and N = 2 (extras: c4_scan split, c5_sharded with the fused exchange)
o=gpurun_out
for n in 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 5 --warmup 3 > $o/r02am_bench_n$n.json 2> $o/r02am_bench_n$n.err; echo "N=$n rc=$?"
  python - $o/r02am_bench_n$n.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")})
for k in ("c4_scan", "c5_sharded"):
    e = d["extra"][k]
    print(k, {q: e.get(q) for q in ("ms_per_step", "updates_per_s", "hbm_roofline_frac_per_gpu", "max_rel_err_vs_unsharded", "efficiency_vs_1gpu_same_kernels", "exchanges_per_step", "halo_aborted", "error")})
PY
done
