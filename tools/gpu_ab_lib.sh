#!/usr/bin/env bash
# A/B of library builds on ONE box: B200_ATTN_LIB selects the build (scratchpad_b200/_lib.py), bench.py is the same.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab_lib.sh TAG'
# Variants (built here, no GPU needed, by
#   B200_OUT=$PWD/_ab/lib_noreduce_oneloop.so B200_OBJDIR=$PWD/_ab/o1 B200_EXTRA_FLAGS="-DB200_NO_CTA_REDUCE -DB200_ONE_LOOP" bash scratchpad_b200/csrc/build.sh
#   B200_OUT=$PWD/_ab/lib_reduce_oneloop.so B200_OBJDIR=$PWD/_ab/o2 B200_EXTRA_FLAGS="-DB200_ONE_LOOP" bash scratchpad_b200/csrc/build.sh ):
#   A = _ab/lib_noreduce_oneloop.so  (round-1 behaviour: one partial per sub-slice, one two-producer loop)
#   B = _ab/lib_reduce_oneloop.so    (in-CTA sub-slice reduction only)
#   C = the in-tree library          (reduction + two loop bodies)
tag="${1:-ab}"; out=$PWD/gpurun_out; mkdir -p $out
A=$PWD/_ab/lib_noreduce_oneloop.so; B=$PWD/_ab/lib_reduce_oneloop.so; C=$PWD/scratchpad_b200/csrc/libb200attn.so
python -m pytest tests/test_paged_gpu.py tests/test_decode_gpu.py -x -q -m gpu > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
run() { lib="$1"; name="$2"; shift 2
  B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py "$@" --no-cpu-baseline --no-e2e --suite off --steps 512 --warmup 16 \
    > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for rep in 1 2; do
  for v in A B C; do
    eval lib=\$$v
    run $lib ${v}_c2_tp8_$rep --workload C2 --shard 8
    run $lib ${v}_c3u_tp8_$rep --workload C3U --shard 8
  done
done
for v in A C; do
  eval lib=\$$v
  run $lib ${v}_c5_tp8 --workload C5 --shard 8
  run $lib ${v}_c3b_tp8 --workload C3B --shard 8
  run $lib ${v}_c2_tp4 --workload C2 --shard 4
done
# tile fetch: plain kernel (page-size 0) vs TMA-capable kernel (page-size -1 = PAGE_AUTO) on random and contiguous slots
for v in B C; do
  eval lib=\$$v
  for w in C2 C3B; do
    run $lib ${v}_${w}_rand_plain --workload $w --page-size 0
    run $lib ${v}_${w}_rand_auto --workload $w --page-size -1
    run $lib ${v}_${w}_contig_plain --workload $w --page-size 0 --index-mode contiguous
    run $lib ${v}_${w}_contig_auto --workload $w --page-size -1 --index-mode contiguous
  done
  run $lib ${v}_C3B_ps16 --workload C3B --page-size 16
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_[ABC]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{f.split('/')[-1][len(tag)+1:-5]:28s} step {d['ms_per_step']*1e3:7.1f} us  kernel {d['roofline']['kernel_us_avg']:7.1f} us  "
              f"frac8T {d['roofline']['frac_of_8tbs']:.3f}  {d['runtime']['tile_fetch'][:6]}  parity {d['parity_check']['ok'] if d.get('parity_check') else None}")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 $out/${tag}_err.log 2>/dev/null
