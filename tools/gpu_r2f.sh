#!/usr/bin/env bash
tag="${1:-r2f}"; out=$PWD/gpurun_out; mkdir -p $out
timeout -s KILL 200 python tools/dbg_pair.py 2>&1 | grep "case\|bad" | tail -30
timeout -s KILL 600 python -m pytest tests/test_extend_gpu.py tests/test_fuzz_gpu.py tests/test_paged_gpu.py tests/test_comparators_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -k "extend" > $out/${tag}_pair_tests.log 2>&1
echo "pair tests rc=$?"; tail -5 $out/${tag}_pair_tests.log
for w in C4 C4P; do for fp in tc tc1; do
  timeout -s KILL 200 python bench.py --workload $w --force-path $fp --steps 16 --warmup 3 --suite off --comparators off > $out/${tag}_${w}_$fp.json 2>> $out/${tag}_err.log
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
