#!/usr/bin/env bash
# Development sweep run on the B200 box through gpurun: bench lines for the BASELINE.json configs
# (per-GPU shard shapes emulated with --shard) into gpurun_out/<tag>_*.json.
tag="${1:-dev}"
out=gpurun_out
mkdir -p $out
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run c2 --workload C2
run c2_contig --workload C2 --index-mode contiguous
run c3b --workload C3B
run c3u --workload C3U
run c3 --workload C3
run c3_nocascade --workload C3 --cascade off
run c5_tp8 --workload C5 --shard 8 --steps 512
run c2_tp8 --workload C2 --shard 8
run c3u_tp8 --workload C3U --shard 8
run c3_tp8 --workload C3 --shard 8
run c3u_tp2 --workload C3U --shard 2
run c3u_tp4 --workload C3U --shard 4
run c4 --workload C4
run c4f --workload C4F
run c4s --workload C4S
run c4p --workload C4P
python tools/bench_table.py $out/${tag}_*.json
tail -5 $out/${tag}_err.log 2>/dev/null
