#!/usr/bin/env bash
# decode loop-structure A/B on one box (and the CTA-pair extend kernel's first bounded run).
# Build the variants first (here, no GPU needed):
#   B200_OUT=$PWD/_ab/lib_A.so B200_OBJDIR=$PWD/_ab/o1 B200_EXTRA_FLAGS="-DB200_NO_CTA_REDUCE -DB200_ONE_LOOP -DB200_FLAT_LOOP" bash scratchpad_b200/csrc/build.sh
#   B200_OUT=$PWD/_ab/lib_C.so B200_OBJDIR=$PWD/_ab/o2 B200_EXTRA_FLAGS="-DB200_FLAT_LOOP" bash scratchpad_b200/csrc/build.sh
#   B200_OUT=$PWD/_ab/lib_E.so B200_OBJDIR=$PWD/_ab/o3 B200_EXTRA_FLAGS="-DB200_NO_CTA_REDUCE" bash scratchpad_b200/csrc/build.sh
tag="${1:-r2e}"; out=$PWD/gpurun_out; mkdir -p $out
echo "== pair kernel: one small case" 
timeout -s KILL 180 python -m pytest "tests/test_extend_gpu.py::test_extend_tcgen05_path_vs_oracle" -x -q -m gpu -k "bf16" > $out/${tag}_pair_small.log 2>&1
rc=$?; echo "pair small rc=$rc"; tail -15 $out/${tag}_pair_small.log
if [ $rc -eq 0 ]; then
  timeout -s KILL 400 python -m pytest tests/test_extend_gpu.py tests/test_fuzz_gpu.py tests/test_paged_gpu.py tests/test_comparators_gpu.py -x -q -m gpu -k "extend" > $out/${tag}_pair_tests.log 2>&1
  echo "pair tests rc=$?"; tail -5 $out/${tag}_pair_tests.log
  for fp in tc tc1; do
    timeout -s KILL 200 python bench.py --workload C4 --force-path $fp --steps 16 --warmup 3 --suite off --comparators off > $out/${tag}_c4_$fp.json 2>> $out/${tag}_err.log
    timeout -s KILL 200 python bench.py --workload C4P --force-path $fp --steps 16 --warmup 3 --suite off --comparators off > $out/${tag}_c4p_$fp.json 2>> $out/${tag}_err.log
  done
  python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_c4*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"{d['value']:.0f} TFLOP/s  {d['ms_per_step']*1e3:.0f} us  parity {d['parity_check']}")
    except Exception as e:
        print(f, "ERR", e)
PY
fi
echo "== decode A/B"
A=$PWD/_ab/lib_A.so; C=$PWD/_ab/lib_C.so; E=$PWD/_ab/lib_E.so; D=$PWD/scratchpad_b200/csrc/libb200attn.so
python -m pytest tests/test_decode_gpu.py tests/test_paged_gpu.py -x -q -m gpu -k "not extend" > $out/${tag}_dec_tests.log 2>&1
echo "decode tests rc=$?"; tail -3 $out/${tag}_dec_tests.log
run() { lib="$1"; name="$2"; shift 2
  B200_ATTN_LIB=$lib timeout -s KILL 200 python bench.py "$@" --no-cpu-baseline --no-e2e --suite off --steps 512 --warmup 16 \
    > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for v in A C E D; do
  eval lib=\$$v
  run $lib ${v}_c2_tp8 --workload C2 --shard 8
  run $lib ${v}_c3b_tp8 --workload C3B --shard 8
  run $lib ${v}_c3u_tp8 --workload C3U --shard 8
  run $lib ${v}_c5_tp8 --workload C5 --shard 8
  run $lib ${v}_c2_tp4 --workload C2 --shard 4
  run $lib ${v}_c2 --workload C2 --page-size 0
  run $lib ${v}_c3b --workload C3B --page-size 0
  run $lib ${v}_c3u --workload C3U --page-size 0
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1]
rows = {}
for f in sorted(glob.glob(f"gpurun_out/{tag}_[ACED]_*.json")):
    name = f.split('/')[-1][len(tag)+1:-5]
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        rows.setdefault(name[2:], {})[name[0]] = (d['ms_per_step']*1e3, d['roofline']['kernel_us_avg'], d['parity_check']['ok'] if d.get('parity_check') else None)
    except Exception as e:
        rows.setdefault(name[2:], {})[name[0]] = ("ERR", str(e)[:40], None)
for w, r in rows.items():
    print(f"{w:10s} " + "  ".join(f"{v}: step {r[v][0]:7.1f} kern {r[v][1]:7.1f} {'ok' if r[v][2] else 'BAD'}" if isinstance(r[v][0], float) else f"{v}: ERR" for v in "ACED" if v in r))
PY
tail -5 $out/${tag}_err.log 2>/dev/null
