#!/usr/bin/env bash
# paged parity tests + the TMA path at TP1 and TP8 shapes (after a change to the TMA issue path)
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 420 python -m pytest tests/test_paged_gpu.py -x -q > $out/${tag}_pytest_paged.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_paged.log
tail -3 $out/${tag}_pytest_paged.log
grep -q "rc=0" $out/${tag}_pytest_paged.log || exit 1
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run C2_ps1 --workload C2
run C2_ps16 --workload C2 --page-size 16
run C2_ps64 --workload C2 --page-size 64
run C3B_ps1 --workload C3B
run C3B_ps16 --workload C3B --page-size 16
run C3B_ps64 --workload C3B --page-size 64
run C3U_ps1 --workload C3U
run C3U_ps16 --workload C3U --page-size 16
run C5_tp8_ps128 --workload C5 --shard 8 --kv-layout split --page-size 128
run C5_tp8_ps64 --workload C5 --shard 8 --kv-layout split --page-size 64
python tools/bench_table.py $out/${tag}_C*.json
