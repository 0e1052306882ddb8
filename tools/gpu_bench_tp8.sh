#!/usr/bin/env bash
# TP8 shard shapes on one GPU: split (reference) vs fused K|V pool layout.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
run() { name="$1"; shift; python bench.py "$@" --no-cpu-baseline > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run c2_tp8_split --workload C2 --shard 8 --kv-layout split
run c2_tp8_fused --workload C2 --shard 8 --kv-layout fused
run c3u_tp8_split --workload C3U --shard 8 --kv-layout split
run c3u_tp8_fused --workload C3U --shard 8 --kv-layout fused
run c5_tp8_split --workload C5 --shard 8 --steps 512 --kv-layout split
run c5_tp8_fused --workload C5 --shard 8 --steps 512 --kv-layout fused
run c3b_tp8_split --workload C3B --shard 8 --kv-layout split
run c3b_tp8_fused --workload C3B --shard 8 --kv-layout fused
python tools/bench_table.py $out/${tag}_*.json
tail -5 $out/${tag}_err.log 2>/dev/null
