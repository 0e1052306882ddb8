#!/usr/bin/env bash
# Shard-shape decode benches (per-GPU shapes of a TP8 split on one GPU) + the tests of the kernels they use.
#   gpurun --timeout 900 -- 'bash tools/gpu_shards.sh TAG'
tag="${1:-shards}"
mkdir -p gpurun_out
python -m pytest tests/test_decode_gpu.py tests/test_paged_gpu.py tests/test_fuzz_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
for w in C2 C3B C3U C5; do
  python bench.py --workload $w --shard 8 --steps 256 --warmup 16 --no-cpu-baseline --suite off > gpurun_out/${tag}_${w}_tp8.json 2>> gpurun_out/${tag}_err.log
done
for w in C2 C3B; do
  python bench.py --workload $w --shard 4 --steps 256 --warmup 16 --no-cpu-baseline --suite off > gpurun_out/${tag}_${w}_tp4.json 2>> gpurun_out/${tag}_err.log
  python bench.py --workload $w --shard 2 --steps 256 --warmup 16 --no-cpu-baseline --suite off > gpurun_out/${tag}_${w}_tp2.json 2>> gpurun_out/${tag}_err.log
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "shards"
for f in sorted(glob.glob(f"gpurun_out/{tag}_*_tp*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"step {d['ms_per_step']*1e3:.1f} us  kernel {d['roofline']['kernel_us_avg']:.1f} us  "
              f"frac {d['roofline']['frac']:.3f}  parity {d['parity_check']['ok'] if d.get('parity_check') else None}")
    except Exception as e:
        print(f, "ERR", e)
PY
