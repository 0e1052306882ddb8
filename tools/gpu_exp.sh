#!/usr/bin/env bash
# Where does the row-gather path of the TMA-capable kernel lose its 6 %?  Library variants (B200_ATTN_LIB) against the build.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
run() { name="$1"; lib="$2"; shift 2; B200_ATTN_LIB=$lib timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
L=$PWD/scratchpad_b200/csrc
for wl in C2 C3B; do
run ${wl}_plain "" --workload $wl
run ${wl}_auto "" --workload $wl --page-size -1
for v in $L/libb200attn_exp*.so; do n=$(basename $v .so); run ${wl}_${n#libb200attn_} $v --workload $wl --page-size -1; done
done
python tools/bench_table.py $out/${tag}_C*.json
tail -5 $out/${tag}_err.log 2>/dev/null
