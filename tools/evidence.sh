#!/bin/bash
# round evidence on the GPU box: tools/evidence.sh TAG   -> gpurun_out/TAG_*  (launch list, ncu captures, bench lines)
tag=$1
o=gpurun_out
FP64M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --time-steps 200 --no-cpu-baseline --no-parity --no-extras > $o/${tag}_launches.log 2>&1
ncu --set full --metrics $FP64M --clock-control none --import-source on -k regex:"k_slab|k_unit" -s 24 -c 2 -f -o $o/${tag}_prof python bench.py --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline --no-parity --no-extras > $o/${tag}_prof.log 2>&1
ncu --set full --metrics $FP64M --cache-control none --clock-control none --import-source on -k regex:"k_slab|k_unit" -s 24 -c 2 -f -o $o/${tag}_prof_warm python bench.py --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline --no-parity --no-extras > $o/${tag}_prof_warm.log 2>&1
ncu --set full --metrics $FP64M --cache-control none --clock-control none -k regex:k_unit -s 12 -c 1 -f -o $o/${tag}_prof_len python bench.py --workload c3_len --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline --no-parity > $o/${tag}_prof_len.log 2>&1
ncu --set full --metrics $FP64M --clock-control none -k regex:k_len_ens -s 3 -c 1 -f -o $o/${tag}_prof_c4 python bench.py --workload c4_len_ensemble --steps 1 --warmup 1 --time-steps 6 --no-cpu-baseline --no-parity > $o/${tag}_prof_c4.log 2>&1
ncu --set full --cache-control none --clock-control none -k regex:k_slab -s 60 -c 1 -f -o $o/${tag}_prof_slab_obs python tools/obs_probe.py VEL 24 > $o/${tag}_prof_slab_obs.log 2>&1
ncu --set full --metrics $FP64M --clock-control none -k regex:k_unit -s 3 -c 1 -f -o $o/${tag}_prof_c2 python bench.py --workload c2_line_ensemble --steps 1 --warmup 1 --time-steps 6 --no-cpu-baseline --no-parity > $o/${tag}_prof_c2.log 2>&1
ncu --set full --metrics $FP64M --clock-control none -k regex:k_adi -s 8 -c 2 -f -o $o/${tag}_prof_adi python bench.py --workload c3_adi --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline --no-parity > $o/${tag}_prof_adi.log 2>&1
ncu --set full --metrics $FP64M --clock-control none -k regex:k_unit -s 6 -c 1 -f -o $o/${tag}_prof_c5 python tools/c5_probe.py 512 8 > $o/${tag}_prof_c5.log 2>&1
# summaries are made here, on the box (gpurun copies back at most 64 MiB): stall / utilisation tables, DRAM traffic + FP64 instruction
# counts (-> profiles/ncu_traffic.json via tools/ncu_traffic.py), dynamic op mix, hottest source lines; then the reports are dropped
for r in prof prof_warm prof_len prof_c4 prof_slab_obs prof_c2 prof_adi prof_c5; do
  python tools/ncu_stalls.py $o/${tag}_$r.ncu-rep > $o/${tag}_$r.stalls.txt 2>&1
  python tools/ncu_opmix.py $o/${tag}_$r.ncu-rep > $o/${tag}_$r.opmix.txt 2>&1
done
cp profiles/ncu_traffic.json $o/${tag}_ncu_traffic.json
python tools/ncu_traffic.py c3_vel $o/${tag}_prof.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py c3_vel_warm $o/${tag}_prof_warm.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py c3_len_warm $o/${tag}_prof_len.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py c4_len_ensemble $o/${tag}_prof_c4.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py c2_line_ensemble $o/${tag}_prof_c2.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py c3_adi $o/${tag}_prof_adi.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_traffic.py c5_one_l_block_16384x512 $o/${tag}_prof_c5.ncu-rep $o/${tag}_ncu_traffic.json > /dev/null 2>&1
python tools/ncu_lines.py $o/${tag}_prof_warm.ncu-rep k_unitILi4ELi3ELi512ELb0 "k_unit<(int)4, (int)3, (int)512, (bool)0>" 40 > $o/${tag}_lines_h2_cn_h2.txt 2>&1
python tools/ncu_lines.py $o/${tag}_prof_warm.ncu-rep k_slabILb0 "k_slab<(bool)0>" 40 > $o/${tag}_lines_slab.txt 2>&1
rm -f $o/${tag}_prof.ncu-rep $o/${tag}_prof_len.ncu-rep $o/${tag}_prof_c4.ncu-rep $o/${tag}_prof_slab_obs.ncu-rep $o/${tag}_prof_c2.ncu-rep $o/${tag}_prof_adi.ncu-rep $o/${tag}_prof_c5.ncu-rep
python bench.py > $o/${tag}_bench_c3_vel.json 2> $o/${tag}_bench_c3_vel.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_c3_vel_reference.json 2>/dev/null
python bench.py --workload c3_len --no-cpu-baseline > $o/${tag}_bench_c3_len.json 2>/dev/null
python bench.py --workload c1_len --no-cpu-baseline > $o/${tag}_bench_c1_len.json 2>/dev/null
python bench.py --workload c3_adi --steps 5 --time-steps 1000 > $o/${tag}_bench_c3_adi.json 2>/dev/null
python bench.py --workload c4_len_ensemble --steps 3 --time-steps 400 --no-cpu-baseline > $o/${tag}_bench_c4_len_ensemble.json 2>/dev/null
python bench.py --workload c2_line_ensemble --steps 3 --time-steps 300 --no-cpu-baseline > $o/${tag}_bench_c2_line_ensemble.json 2>/dev/null
python tools/c5_probe.py 4096 40 > $o/${tag}_c5_one_gpu.log 2>&1
python tools/obs_probe.py VEL 512 > $o/${tag}_obs_every_step.log 2>&1
python tools/obs_probe.py LEN 512 >> $o/${tag}_obs_every_step.log 2>&1
ION_NO_FUSED_OBS=1 python tools/obs_probe.py VEL 512 >> $o/${tag}_obs_every_step.log 2>&1
ION_NO_FUSED_OBS=1 python tools/obs_probe.py LEN 512 >> $o/${tag}_obs_every_step.log 2>&1
for f in $o/${tag}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value %.3e' % d['value'], 'us/step', d.get('us_per_time_step'), 'frac', d.get('hbm_roofline_frac_step'), 'parity', (d.get('parity') or {}).get('max_rel_err'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
cat $o/${tag}_obs_every_step.log
