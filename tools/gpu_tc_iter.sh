#!/usr/bin/env bash
# One GPU iteration on the tcgen05 extend kernel: parity tests, C4 bench, optional timeline trace.
# usage: gpu_tc_iter.sh <tag> [trace_rows]
tag="${1:-dev}"; rows="${2:-0}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 90 python -m pytest tests/test_extend_gpu.py -m gpu -q -x --timeout 40 2>&1 | tail -3
timeout -s KILL 60 python bench.py --workload C4 > $out/${tag}_c4.json 2> $out/${tag}_c4.err; tail -2 $out/${tag}_c4.err
python tools/bench_table.py $out/${tag}_c4.json
if [ "$rows" != "0" ]; then
  timeout -s KILL 60 python tools/tc_trace.py $rows > $out/${tag}_trace.txt 2>&1; tail -n $((rows + 2)) $out/${tag}_trace.txt | cut -c1-260
fi
