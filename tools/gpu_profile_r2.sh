#!/usr/bin/env bash
# Round-2 ncu evidence for profiles/: launch list of the default bench command + --set full captures of the dominant
# kernels.  Each ncu run only after the same command exited 0 without ncu.  Run through gpurun (1 GPU).
tag="${1:-r2p}"
out=gpurun_out
mkdir -p $out
K='group_kernel|lcp_kernel|plan_kernel|decode_mma|decode_paged|merge_kernel|kv_store|extend_tc|extend_mma'
BASE="--suite off --comparators off --no-e2e --no-cpu-baseline --no-parity"
# 1. launch list of the default command's headline leg (CUDA-graph mode as the driver runs it; eager as a fallback)
CMD="python bench.py --steps 20 --warmup 5 $BASE"
timeout -s KILL 120 $CMD > $out/${tag}_plain.log 2>&1 && timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $out/${tag}_launches_c3b.csv $CMD > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $out/${tag}_launches_c3b.csv)"
if [ "$(wc -l < $out/${tag}_launches_c3b.csv)" -lt 20 ]; then
  timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file $out/${tag}_launches_c3b.csv $CMD --mode eager > $out/${tag}_ncu_launches.log 2>&1
fi
cap() { name="$1"; kre="$2"; skip="$3"; shift 3
  timeout -s KILL 120 python bench.py "$@" $BASE --mode eager --steps 8 --warmup 3 > $out/${tag}_plain_$name.log 2>&1 || { echo "plain $name failed"; return; }
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c 1 -o $out/${tag}_prof_$name python bench.py "$@" $BASE --mode eager --steps 8 --warmup 3 > $out/${tag}_ncu_$name.log 2>&1
  echo "capture $name rc=$?"
  # summarise on the box (gpurun copies at most 64 MiB back): keep the .ncu-rep of the headline and the extend kernel only
  python tools/ncu_summary.py $out/${tag}_prof_$name.ncu-rep > $out/${tag}_ncu_$name.txt 2>&1
  case "$name" in c3b|c4) ;; *) rm -f $out/${tag}_prof_$name.ncu-rep ;; esac; }
cap c3b 'decode_(mma|paged)' 6 --layers 4
cap c3b_contig 'decode_(mma|paged)' 6 --layers 4 --index-mode contiguous
cap c2 'decode_(mma|paged)' 6 --workload C2 --layers 4
cap c3u 'decode_(mma|paged)' 6 --workload C3U --layers 4
cap c2_tp8 'decode_(mma|paged)' 6 --workload C2 --shard 8 --layers 4
cap c5_tp8 'decode_(mma|paged)' 6 --workload C5 --shard 8 --layers 4
cap c3b_tp2 'decode_(mma|paged)' 6 --shard 2 --layers 4
cap c3b_tp4 'decode_(mma|paged)' 6 --shard 4 --layers 4
cap c3b_tp8 'decode_(mma|paged)' 6 --shard 8 --layers 4
cap c4 'extend_tc' 2 --workload C4 --layers 2
cap c4_contig 'extend_tc' 2 --workload C4 --layers 2 --index-mode contiguous
cap c4_pair 'extend_tc2' 2 --workload C4 --layers 2 --force-path tc2
python tools/ncu_summary.py --launches $out/${tag}_launches_c3b.csv > $out/${tag}_launches_c3b.txt 2>&1
ls -la $out/ | head -40
