#!/usr/bin/env bash
tag="${1:-r2l}"; out=$PWD/gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_extend_gpu.py tests/test_fuzz_gpu.py tests/test_paged_gpu.py tests/test_decode_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log
N=$PWD/_ab/lib_noopt.so; O=$PWD/scratchpad_b200/csrc/libb200attn.so
for rep in 1 2; do for v in N O; do eval lib=\$$v
  for w in C4 C4F C4S C4P; do
    B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py --workload $w --steps 16 --warmup 3 --suite off --comparators off --no-e2e --no-parity > $out/${tag}_${v}_${w}_$rep.json 2>> $out/${tag}_err.log
  done
  B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py --workload C3 --steps 256 --warmup 16 --suite off --comparators off --no-e2e --no-cpu-baseline > $out/${tag}_${v}_C3_$rep.json 2>> $out/${tag}_err.log
done; done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_[NO]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][len(sys.argv[1])+1:-5], f"{d['value']:.0f} {d['unit'][:8]}  {d['ms_per_step']*1e3:.1f} us")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $out/${tag}_err.log
