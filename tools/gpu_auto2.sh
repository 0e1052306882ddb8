#!/usr/bin/env bash
# short form of gpu_auto.sh: paged parity tests + C2/C3B random & contiguous, off / auto / page 16
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 420 python -m pytest tests/test_paged_gpu.py -x -q > $out/${tag}_pytest_paged.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_paged.log
tail -3 $out/${tag}_pytest_paged.log
grep -q "rc=0" $out/${tag}_pytest_paged.log || exit 1
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for wl in C2 C3B; do
  run ${wl}_random_off --workload $wl
  run ${wl}_random_auto --workload $wl --page-size -1
  run ${wl}_contig_auto --workload $wl --index-mode contiguous --page-size -1
  run ${wl}_ps16 --workload $wl --page-size 16
done
python tools/bench_table.py $out/${tag}_C*.json
