#!/usr/bin/env bash
# TP8-shard shapes (1 kv head per GPU): row gather (split / fused K|V rows) against paged pools (16 / 64 / 128-token pages).
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for wl in C5 C3U C2; do
  run ${wl}_tp8_ps1_split --workload $wl --shard 8 --kv-layout split
  run ${wl}_tp8_ps1_fused --workload $wl --shard 8 --kv-layout fused
  run ${wl}_tp8_ps16_split --workload $wl --shard 8 --kv-layout split --page-size 16
  run ${wl}_tp8_ps64_split --workload $wl --shard 8 --kv-layout split --page-size 64
  run ${wl}_tp8_ps128_split --workload $wl --shard 8 --kv-layout split --page-size 128
  run ${wl}_tp8_ps128_fused --workload $wl --shard 8 --kv-layout fused --page-size 128
done
python tools/bench_table.py $out/${tag}_C*.json
tail -5 $out/${tag}_err.log 2>/dev/null
