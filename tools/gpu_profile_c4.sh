#!/usr/bin/env bash
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
CMD3="python bench.py --workload C4 --steps 4 --warmup 3 --layers 2"
timeout -s KILL 120 $CMD3 > $out/${tag}_plain3.log 2>&1 && timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:extend_tc -s 2 -c 1 -o $out/${tag}_prof_c4 $CMD3 > $out/${tag}_ncu_c4.log 2>&1
tail -n 2 $out/${tag}_ncu_c4.log
