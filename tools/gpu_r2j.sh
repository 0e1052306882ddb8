#!/usr/bin/env bash
tag="${1:-r2j}"; out=$PWD/gpurun_out; mkdir -p $out
B200_ATTN_LIB=$PWD/_ab/lib_tc2_trace.so timeout 200 python tools/tc2_trace.py 2>&1 | tail -16
timeout -s KILL 200 python tools/dbg_pair.py 2>&1 | grep "path=1" | grep -v "bad rows 0 " 
for w in C4 C4P; do for fp in tc tc1; do
  timeout -s KILL 200 python bench.py --workload $w --force-path $fp --steps 16 --warmup 3 --suite off --comparators off --no-e2e > $out/${tag}_${w}_$fp.json 2>> $out/${tag}_err.log
done; done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_C4*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['value']:.0f} TFLOP/s  {d['ms_per_step']*1e3:.0f} us  parity {d['parity_check']['ok']} {d['parity_check']['vs_cpu_oracle']['max_abs']:.2g}")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $out/${tag}_err.log
