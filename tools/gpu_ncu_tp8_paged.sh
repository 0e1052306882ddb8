#!/usr/bin/env bash
# ncu of the per-page-run TMA path (one kv head per CTA, 128-token tiles, 64-token pages -> 8 KB boxes) at the C5/TP8 shard,
# and of the whole-tile path at 128-token pages.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
for ps in 64 128; do
CMD="python bench.py --workload C5 --shard 8 --kv-layout split --page-size $ps --steps 16 --warmup 4 --no-e2e --no-cpu-baseline --mode eager --layers 2"
timeout -s KILL 120 $CMD > $out/${tag}_plain_ps$ps.log 2>&1 && timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:decode_paged -s 6 -c 1 -o $out/${tag}_prof_c5tp8_ps$ps $CMD > $out/${tag}_ncu_ps$ps.log 2>&1
tail -n 1 $out/${tag}_ncu_ps$ps.log
done
