#!/usr/bin/env bash
# A/B on one box: the tree at _ab_old (git worktree of an earlier commit, built in place) against the current tree.
tag="${1:-dev}"; out=$PWD/gpurun_out; mkdir -p $out
run() { dir="$1"; name="$2"; shift 2; (cd $dir && timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log); }
for rep in 1 2; do
run _ab_old old_c5_fused_$rep --workload C5 --shard 8 --kv-layout fused
run . new_c5_fused_$rep --workload C5 --shard 8 --kv-layout fused
run _ab_old old_c3u_tp8_$rep --workload C3U --shard 8
run . new_c3u_tp8_$rep --workload C3U --shard 8
done
run _ab_old old_c2_tp8 --workload C2 --shard 8
run . new_c2_tp8 --workload C2 --shard 8
python tools/bench_table.py $out/${tag}_old_*.json $out/${tag}_new_*.json
tail -5 $out/${tag}_err.log 2>/dev/null
