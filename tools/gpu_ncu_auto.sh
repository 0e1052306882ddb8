#!/usr/bin/env bash
# ncu capture of the TMA-capable decode kernel when every tile falls back to the row gather (random slots, PAGE_AUTO).
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
CMD="python bench.py --page-size -1 --steps 32 --warmup 8 --no-e2e --no-cpu-baseline --mode eager --layers 4"
timeout -s KILL 120 $CMD > $out/${tag}_plain.log 2>&1 && timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:decode_paged -s 12 -c 1 -o $out/${tag}_prof_c2_auto $CMD > $out/${tag}_ncu.log 2>&1
tail -n 2 $out/${tag}_ncu.log
