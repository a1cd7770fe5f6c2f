#!/bin/bash
# c3 (BASELINE configs[2]) under ncu: every movement launch of one tick (per-launch metrics) and a full capture of the iter-0 window kernel
O=gpurun_out
C3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --e2e-steps 0"
$C3 > $O/r02_c3_cmd.json 2> /dev/null || exit 1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
# tick 3 = launches after 3 warm-up ticks of 28 launches each (4 pre-pass + 24 movement), plus one k_temperature at tick 2
ncu --metrics $M --clock-control none -k regex:"k_substep|k_fire_prepass" -s 84 -c 28 --csv --log-file $O/r02_c3_one_tick_16384.csv $C3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_substep_win -s 72 -c 1 -f -o $O/r02_k1_gen $C3 > /dev/null 2>&1
ls -la $O/r02_c3* $O/r02_k1_gen*
