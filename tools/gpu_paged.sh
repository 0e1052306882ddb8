#!/usr/bin/env bash
# Paged-pool decode (TMA page copies): parity tests, then page size 1 vs 16 vs 64 on the shard shapes where the row is small.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 420 python -m pytest tests/test_paged_gpu.py -x -q > $out/${tag}_pytest_paged.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_paged.log
tail -15 $out/${tag}_pytest_paged.log
grep -q "rc=0" $out/${tag}_pytest_paged.log || exit 1
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for ps in 1 16 64; do
  run c5_tp8_ps$ps --workload C5 --shard 8 --steps 512 --page-size $ps --kv-layout split
  run c2_ps$ps --workload C2 --page-size $ps
  run c3u_tp8_ps$ps --workload C3U --shard 8 --page-size $ps --kv-layout split
done
run c3b_ps1 --workload C3B
run c3b_ps16 --workload C3B --page-size 16
python tools/bench_table.py $out/${tag}_*_ps*.json
tail -5 $out/${tag}_err.log 2>/dev/null
