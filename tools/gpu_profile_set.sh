#!/usr/bin/env bash
# ncu evidence for profiles/: launch list of the default bench command + one --set full capture of the
# dominant decode kernel, of the tcgen05 extend kernel (C4) and of its shared-prefix mode (C3 cascade).
# Each ncu run only after the same command exited 0 without ncu.  Run through gpurun (1 GPU).
tag="${1:-dev}"
out=gpurun_out
mkdir -p $out
K='group_kernel|lcp_kernel|plan_kernel|decode_mma|merge_kernel|kv_store|extend_tc|extend_mma'
CMD="python bench.py --steps 32 --warmup 8 --no-e2e --no-cpu-baseline --mode eager --layers 4"
timeout -s KILL 120 $CMD > $out/${tag}_plain.log 2>&1 && timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $out/${tag}_launches_c2.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:decode_mma -s 12 -c 2 -o $out/${tag}_prof_c2 $CMD > $out/${tag}_ncu_c2.log 2>&1
CMD2="python bench.py --workload C3 --steps 16 --warmup 4 --no-e2e --no-cpu-baseline --mode eager --layers 2"
timeout -s KILL 120 $CMD2 > $out/${tag}_plain2.log 2>&1 && timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 120 --csv --log-file $out/${tag}_launches_c3.csv $CMD2 > $out/${tag}_ncu_launches3.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:extend_tc -s 6 -c 1 -o $out/${tag}_prof_c3cascade $CMD2 > $out/${tag}_ncu_c3.log 2>&1
CMD3="python bench.py --workload C4 --steps 4 --warmup 3 --layers 2"
timeout -s KILL 120 $CMD3 > $out/${tag}_plain3.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:extend_tc -s 2 -c 1 -o $out/${tag}_prof_c4 $CMD3 > $out/${tag}_ncu_c4.log 2>&1
tail -n 2 $out/${tag}_ncu_c2.log $out/${tag}_ncu_c3.log $out/${tag}_ncu_c4.log
