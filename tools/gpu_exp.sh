#!/usr/bin/env bash
# Where does the row-gather path of the TMA-capable kernel lose its 6 %?  Library variants built with -DB200_PAGED_EXPERIMENT.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
run() { name="$1"; lib="$2"; shift 2; B200_ATTN_LIB=$lib timeout -s KILL 150 python bench.py "$@" --no-cpu-baseline --no-e2e --steps 1024 > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
L=$PWD/scratchpad_b200/csrc
run c2_plain "" --workload C2
run c2_auto "" --workload C2 --page-size -1
run c2_exp1 $L/libb200attn_exp1.so --workload C2 --page-size -1
run c2_exp2 $L/libb200attn_exp2.so --workload C2 --page-size -1
run c3b_plain "" --workload C3B
run c3b_auto "" --workload C3B --page-size -1
run c3b_exp1 $L/libb200attn_exp1.so --workload C3B --page-size -1
run c3b_exp2 $L/libb200attn_exp2.so --workload C3B --page-size -1
python tools/bench_table.py $out/${tag}_c*.json
tail -5 $out/${tag}_err.log 2>/dev/null
