#!/bin/bash
# the round's records: default bench line, reference arm, small worlds, ncu launch list / per-launch metrics / full captures
# (each ncu pass only after the same command has exited 0 without ncu); everything lands in gpurun_out/
O=gpurun_out
python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err || echo "bench default failed"
python bench.py --impl reference > $O/r02_bench_ref.json 2> /dev/null || echo "bench ref failed"
python tests/perf/bench_small.py > $O/r02_small_worlds.json 2> /dev/null || echo "bench small failed"
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu --e2e-steps 0"
$CMD > $O/r02_ncu_cmd.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c2_16384.csv $CMD > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:k_substep_win -s 72 -c 24 --csv --log-file $O/r02_k1_one_tick_c2_16384.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_substep_win -s 72 -c 1 -f -o $O/r02_k1_sand $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_substep_win -s 80 -c 1 -f -o $O/r02_k1_liq $CMD > /dev/null 2>&1
C3="python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu --e2e-steps 0"
$C3 > $O/r02_bench_c3.json 2> /dev/null && ncu --set full --clock-control none --import-source on -k regex:k_temperature -s 1 -c 1 -f -o $O/r02_k2 $C3 > /dev/null 2>&1
ls -la $O/r02_*
