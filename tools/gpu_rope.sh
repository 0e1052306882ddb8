#!/usr/bin/env bash
# Rope parity tests + the cost of the rotary embedding in a decode layer-call: none / separate launch / fused with the KV write.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 300 python -m pytest tests/test_rope_gpu.py -x -q > $out/${tag}_pytest_rope.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_rope.log
tail -4 $out/${tag}_pytest_rope.log
run() { name="$1"; shift; timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for r in off separate fused; do
  run c2_rope_$r --workload C2 --rope $r
  run c2_tp8_rope_$r --workload C2 --shard 8 --rope $r
done
python tools/bench_table.py $out/${tag}_*rope*.json
tail -5 $out/${tag}_err.log 2>/dev/null
