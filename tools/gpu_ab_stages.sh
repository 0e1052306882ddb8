#!/usr/bin/env bash
# ring depth / CTAs per SM of the split-KV kernels, A/B on ONE box.  Variants (built here, no GPU needed):
#   for s in 4 5 7; do mkdir -p _ab/o_s$s; B200_OUT=$PWD/_ab/lib_s$s.so B200_OBJDIR=$PWD/_ab/o_s$s \
#     B200_EXTRA_FLAGS="-DB200_DEC_STAGES=$s -DB200_DEC_MINBLOCKS=1" bash scratchpad_b200/csrc/build.sh; done
#   base = the in-tree library (3 stages of 32 KiB, 2 CTAs per SM)
tag="${1:-r5c}"; out=$PWD/gpurun_out; mkdir -p $out
run() { lib="$1"; name="$2"; shift 2
  B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py "$@" --no-cpu-baseline --no-e2e --suite off --comparators off --steps 128 --warmup 16 \
    > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
base=$PWD/scratchpad_b200/csrc/libb200attn.so
for v in base s7 s5 s4; do
  lib=$PWD/_ab/lib_$v.so; [ $v = base ] && lib=$base
  run $lib ${v}_C3B_ps64 --workload C3B --page-size 64
  [ $v = s4 ] && continue
  run $lib ${v}_C3B_rand --workload C3B
  run $lib ${v}_C3U_rand --workload C3U
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_*_C3*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{f.split('/')[-1][len(tag)+1:-5]:20s} step {d['ms_per_step']*1e3:7.1f} us  kernel {d['roofline']['kernel_us_avg']:7.1f} us  "
              f"frac8T {d['roofline']['frac_of_8tbs']:.3f}  grid {d['runtime'].get('grid')}  parity {d['parity_check']['ok'] if d.get('parity_check') else None}")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 $out/${tag}_err.log 2>/dev/null
