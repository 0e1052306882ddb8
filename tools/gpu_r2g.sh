#!/usr/bin/env bash
tag="${1:-r2g}"; out=$PWD/gpurun_out; mkdir -p $out
timeout -s KILL 200 python tools/dbg_pair.py 2>&1 | grep "path=1" | tail -12
for v in l5s10 l11s12 default; do
  lib=$PWD/_ab/lib_tc2_$v.so; [ $v = default ] && lib=$PWD/scratchpad_b200/csrc/libb200attn.so
  for w in C4 C4P; do
    B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py --workload $w --force-path tc --steps 16 --warmup 3 --suite off --comparators off --no-e2e > $out/${tag}_${w}_$v.json 2>> $out/${tag}_err.log
  done
done
timeout -s KILL 200 python bench.py --workload C4 --force-path tc1 --steps 16 --warmup 3 --suite off --comparators off --no-e2e > $out/${tag}_C4_tc1.json 2>> $out/${tag}_err.log
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_C4*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['value']:.0f} TFLOP/s  {d['ms_per_step']*1e3:.0f} us  parity {d['parity_check']['ok']}")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $out/${tag}_err.log
