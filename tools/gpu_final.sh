#!/usr/bin/env bash
# Round-end check on one box: full GPU suite, smoke(), the default bench line (both arms), refreshed C2 ncu evidence.
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -4 $out/${tag}_pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
tail -7 $out/${tag}_smoke.log
timeout -s KILL 300 python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_default.err; echo "ref rc=$?"
python tools/bench_table.py $out/${tag}_bench_default.json
K='group_kernel|lcp_kernel|plan_kernel|decode_mma|decode_paged|merge_kernel|kv_store|extend_tc|extend_mma'
CMD="python bench.py --steps 32 --warmup 8 --no-e2e --no-cpu-baseline --mode eager --layers 4"
timeout -s KILL 120 $CMD > $out/${tag}_plain.log 2>&1 && timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $out/${tag}_launches_c2.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:decode_mma -s 12 -c 2 -o $out/${tag}_prof_c2 $CMD > $out/${tag}_ncu_c2.log 2>&1
tail -n 2 $out/${tag}_ncu_c2.log
