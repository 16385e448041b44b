#!/bin/bash
# round evidence on the GPU box: tools/evidence.sh TAG   -> gpurun_out/TAG_*  (launch list, ncu captures, bench lines)
tag=$1
o=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --time-steps 200 --no-cpu-baseline > $o/${tag}_launches.log 2>&1
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum --clock-control none --import-source on -k regex:"k_slab|k_unit" -s 24 -c 2 -f -o $o/${tag}_prof python bench.py --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline > $o/${tag}_prof.log 2>&1
ncu --set full --cache-control none --clock-control none --import-source on -k regex:"k_slab|k_unit" -s 24 -c 2 -f -o $o/${tag}_prof_warm python bench.py --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline > $o/${tag}_prof_warm.log 2>&1
ncu --set full --clock-control none -k regex:k_unit -s 12 -c 1 -f -o $o/${tag}_prof_c4 python bench.py --workload c4_len_ensemble --steps 1 --warmup 1 --time-steps 6 --no-cpu-baseline > $o/${tag}_prof_c4.log 2>&1
ncu --set full --clock-control none -k regex:"k_adi_l|k_unit" -s 12 -c 2 -f -o $o/${tag}_prof_adi python bench.py --workload c3_adi --steps 1 --warmup 1 --time-steps 10 --no-cpu-baseline > $o/${tag}_prof_adi.log 2>&1
python bench.py > $o/${tag}_bench_c3_vel.json 2> $o/${tag}_bench_c3_vel.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_c3_vel_reference.json 2>/dev/null
python bench.py --workload c3_len --no-cpu-baseline > $o/${tag}_bench_c3_len.json 2>/dev/null
python bench.py --workload c1_len --no-cpu-baseline > $o/${tag}_bench_c1_len.json 2>/dev/null
python bench.py --workload c3_adi --steps 5 --time-steps 1000 > $o/${tag}_bench_c3_adi.json 2>/dev/null
python bench.py --workload c4_len_ensemble --steps 3 --time-steps 400 --no-cpu-baseline > $o/${tag}_bench_c4_len_ensemble.json 2>/dev/null
python bench.py --workload c2_line_ensemble --steps 3 --time-steps 300 --no-cpu-baseline > $o/${tag}_bench_c2_line_ensemble.json 2>/dev/null
for f in $o/${tag}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value %.3e' % d['value'], 'us/step', d.get('us_per_time_step'), 'frac', d.get('hbm_roofline_frac_step'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
