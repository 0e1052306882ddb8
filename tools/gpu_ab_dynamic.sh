#!/usr/bin/env bash
# DESIGN 3.2 "Dynamic tail": equal static tile ranges (A = -DB200_STATIC_TILES build) vs 7/8 static + dynamically claimed
# chunks (B = the in-tree library) on ONE box.  Build A first (here, no GPU needed):
#   mkdir -p _ab/o_static; B200_OUT=$PWD/_ab/lib_static.so B200_OBJDIR=$PWD/_ab/o_static B200_EXTRA_FLAGS="-DB200_STATIC_TILES" bash scratchpad_b200/csrc/build.sh
tag="${1:-r5e}"; out=$PWD/gpurun_out; mkdir -p $out
run() { lib="$1"; name="$2"; shift 2
  B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py "$@" --no-cpu-baseline --no-e2e --suite off --comparators off --steps 128 --warmup 16 \
    > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
A=$PWD/_ab/lib_static.so; B=$PWD/scratchpad_b200/csrc/libb200attn.so
for v in A B; do
  eval lib=\$$v
  run $lib ${v}_C3B_ps64 --workload C3B --page-size 64
  run $lib ${v}_C3B_contig --workload C3B --index-mode contiguous
  run $lib ${v}_C3U_contig --workload C3U --index-mode contiguous
  run $lib ${v}_C2_contig --workload C2 --index-mode contiguous
done
# scattered slots: the plan's probe keeps the static ranges, B must read like A
run $B B_C3B_rand --workload C3B
run $B B_C3U_rand --workload C3U
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_[AB]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{f.split('/')[-1][len(tag)+1:-5]:22s} step {d['ms_per_step']*1e3:7.1f} us  kernel {d['roofline']['kernel_us_avg']:7.1f} us  "
              f"frac8T {d['roofline']['frac_of_8tbs']:.3f}  parity {d['parity_check']['ok'] if d.get('parity_check') else None}")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 $out/${tag}_err.log 2>/dev/null
