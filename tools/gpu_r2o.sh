#!/usr/bin/env bash
tag="${1:-r2o}"; out=$PWD/gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_extend_gpu.py tests/test_fuzz_gpu.py tests/test_paged_gpu.py -x -q -m gpu -k "extend or fuzz" > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $out/${tag}_tests.log
P=$PWD/scratchpad_b200/csrc/libb200attn.so; N=$PWD/_ab/lib_noprefetch.so
for rep in 1 2; do for v in P N; do eval lib=\$$v
  for w in C4 C4S C4F; do
    B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py --workload $w --steps 16 --warmup 3 --suite off --comparators off --no-e2e --no-parity > $out/${tag}_${v}_${w}_$rep.json 2>> $out/${tag}_err.log
  done
  B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py --workload C3 --steps 256 --warmup 16 --suite off --comparators off --no-e2e --no-cpu-baseline --no-parity > $out/${tag}_${v}_C3_$rep.json 2>> $out/${tag}_err.log
done; done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_[PN]_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][len(sys.argv[1])+1:-5], f"{d['value']:.0f}  {d['ms_per_step']*1e3:.1f} us")
    except Exception as e:
        print(f, "ERR", e)
PY
grep -v -i "warn" $out/${tag}_err.log | tail -3
