#!/usr/bin/env bash
# Paged decode, second sweep: page >= tile on the TP8 shards, fused K|V layout + pages, cascade + pages; then the ncu
# evidence of decode_paged_kernel (launch list + one --set full capture) at the north_star target shape (C3B).
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run c5_tp8_ps128 --workload C5 --shard 8 --page-size 128 --kv-layout split
run c5_tp8_ps128_fused --workload C5 --shard 8 --page-size 128 --kv-layout fused
run c5_tp8_ps64_fused --workload C5 --shard 8 --page-size 64 --kv-layout fused
run c5_tp8_ps1_fused --workload C5 --shard 8 --kv-layout fused
run c3u_tp8_ps128_fused --workload C3U --shard 8 --page-size 128 --kv-layout fused
run c3u_ps16 --workload C3U --page-size 16
run c3u_ps1 --workload C3U
run c3b_ps64 --workload C3B --page-size 64
run c3_cascade_ps16 --workload C3 --page-size 16
run c3_cascade_ps1 --workload C3
python tools/bench_table.py $out/${tag}_c*.json
K='plan_kernel|decode_paged|merge_kernel|kv_store'
CMD="python bench.py --workload C3B --page-size 16 --steps 32 --warmup 8 --no-e2e --no-cpu-baseline --mode eager --layers 4"
timeout -s KILL 120 $CMD > $out/${tag}_plain.log 2>&1 && timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 200 --csv --log-file $out/${tag}_launches_c3b_paged.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:decode_paged -s 12 -c 2 -o $out/${tag}_prof_c3b_paged $CMD > $out/${tag}_ncu_c3b.log 2>&1
tail -n 2 $out/${tag}_ncu_c3b.log
tail -5 $out/${tag}_err.log 2>/dev/null
