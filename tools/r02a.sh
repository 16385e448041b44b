#!/bin/bash
# round-2 first GPU pass: full gpu test suite (incl. benchmarked-shape parity), self-verifying bench, FP64 peak, ncu capture with FP64 op counts
o=gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 > $o/r02a_gputests.log 2>&1; echo "pytest rc=$?" | tee -a $o/r02a_gputests.log
tail -15 $o/r02a_gputests.log
python bench.py > $o/r02a_bench_c3_vel.json 2> $o/r02a_bench_c3_vel.err; tail -c 2500 $o/r02a_bench_c3_vel.json
python bench.py --workload c3_len --no-cpu-baseline > $o/r02a_bench_c3_len.json 2>/dev/null; tail -c 1200 $o/r02a_bench_c3_len.json
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum --cache-control none --clock-control none --import-source on -k regex:"k_slab|k_unit" -s 24 -c 2 -f -o $o/r02a_prof_warm python bench.py --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline --no-parity > $o/r02a_prof.log 2>&1
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum --cache-control none --clock-control none -k regex:"k_unit" -s 12 -c 1 -f -o $o/r02a_prof_len python bench.py --workload c3_len --steps 1 --warmup 1 --time-steps 20 --no-cpu-baseline --no-parity > $o/r02a_prof_len.log 2>&1
ls -la $o | tail -8
