#!/usr/bin/env bash
# Round-2 (late) ncu refresh after the tile-schedule change: launch list of the default bench command + --set full captures
# of the split-KV kernel on the headline (scattered slots: static ranges), on consecutive slots and on 64-token pages
# (dynamically claimed tail).  Each ncu run only after the same command exited 0 without ncu.  Run through gpurun (1 GPU).
tag="${1:-r5p}"
out=gpurun_out
mkdir -p $out
K='group_kernel|lcp_kernel|plan_kernel|decode_mma|decode_paged|merge_kernel|kv_store|extend_tc|extend_mma'
BASE="--suite off --comparators off --no-e2e --no-cpu-baseline --no-parity"
CMD="python bench.py --steps 20 --warmup 5 $BASE"
timeout -s KILL 120 $CMD > $out/${tag}_plain.log 2>&1 && timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $out/${tag}_launches_c3b.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $out/${tag}_launches_c3b.csv)"
cap() { name="$1"; kre="$2"; skip="$3"; shift 3
  timeout -s KILL 120 python bench.py "$@" $BASE --mode eager --steps 8 --warmup 3 > $out/${tag}_plain_$name.log 2>&1 || { echo "plain $name failed"; return; }
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c 1 -o $out/${tag}_prof_$name python bench.py "$@" $BASE --mode eager --steps 8 --warmup 3 > $out/${tag}_ncu_$name.log 2>&1
  echo "capture $name rc=$?"
  python tools/ncu_summary.py $out/${tag}_prof_$name.ncu-rep > $out/${tag}_ncu_$name.txt 2>&1
  case "$name" in c3b) ;; *) rm -f $out/${tag}_prof_$name.ncu-rep ;; esac; }
cap c3b 'decode_(mma|paged)' 6 --layers 4
cap c3b_contig 'decode_(mma|paged)' 6 --layers 4 --index-mode contiguous
cap c3b_ps64 'decode_(mma|paged)' 6 --layers 4 --page-size 64
python tools/ncu_summary.py --launches $out/${tag}_launches_c3b.csv > $out/${tag}_launches_c3b.txt 2>&1
cat $out/${tag}_launches_c3b.txt | head -n 30
for n in c3b c3b_contig c3b_ps64; do echo "== $n"; grep -E "gpu__time_duration|dram__bytes_read.sum |dram__bytes_write|dram_throughput|registers_per_thread|grid_size" $out/${tag}_ncu_$n.txt | head -n 8; done
