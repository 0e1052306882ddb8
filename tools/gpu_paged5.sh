#!/usr/bin/env bash
# paged parity tests + the per-page-run TMA path (page < tile) at the TP8 shards, and the whole-tile path as control
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 420 python -m pytest tests/test_paged_gpu.py -x -q > $out/${tag}_pytest_paged.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_paged.log
tail -3 $out/${tag}_pytest_paged.log
grep -q "rc=0" $out/${tag}_pytest_paged.log || exit 1
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run C5_tp8_ps1_fused --workload C5 --shard 8 --kv-layout fused
run C5_tp8_ps32 --workload C5 --shard 8 --kv-layout split --page-size 32
run C5_tp8_ps64 --workload C5 --shard 8 --kv-layout split --page-size 64
run C5_tp8_ps128 --workload C5 --shard 8 --kv-layout split --page-size 128
run C3U_tp8_ps1_fused --workload C3U --shard 8 --kv-layout fused
run C3U_tp8_ps64 --workload C3U --shard 8 --kv-layout split --page-size 64
run C3B_tp8_ps1_fused --workload C3B --shard 8 --kv-layout fused
run C3B_tp8_ps16 --workload C3B --shard 8 --kv-layout split --page-size 16
run C3B_tp8_ps32 --workload C3B --shard 8 --kv-layout split --page-size 32
run C3B_ps16 --workload C3B --page-size 16
python tools/bench_table.py $out/${tag}_C*.json
