#!/usr/bin/env bash
tag="${1:-r2n}"; out=$PWD/gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_extend_gpu.py tests/test_fuzz_gpu.py tests/test_paged_gpu.py tests/test_comparators_gpu.py tests/test_engine_boundary_gpu.py tests/test_decode_gpu.py -x -q -m gpu > $out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $out/${tag}_tests.log
for rep in 1 2; do for v in tma notma; do
  [ $v = notma ] && export B200_EXTEND_NO_TMA=1 || unset B200_EXTEND_NO_TMA
  for w in C4 C4P C4S; do
    for im in random contiguous; do
    timeout -s KILL 200 python bench.py --workload $w --index-mode $im --steps 16 --warmup 3 --suite off --comparators off --no-e2e > $out/${tag}_${v}_${w}_${im}_$rep.json 2>> $out/${tag}_err.log
    done
  done
done; done
unset B200_EXTEND_NO_TMA
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*tma_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][len(sys.argv[1])+1:-5], f"{d['value']:.0f} TFLOP/s  {d['ms_per_step']*1e3:.1f} us  tma tiles {d['runtime']['kv_tiles_fetched_by_tma_per_launch']} of {d['runtime']['kv_tiles_per_launch']}  parity {d['parity_check']['ok']}")
    except Exception as e:
        print(f, "ERR", e)
PY
grep -v -i "warn" $out/${tag}_err.log | tail -3
