#!/usr/bin/env bash
# DESIGN 3.2 "Dynamic tail": sweep of the dynamic fraction / chunk size (B200_SCHED="frac_den,chunk_den"; "0,0" = static
# ranges only) on ONE box, C3B over 64-token pages and over consecutive slots of the token-granular pool.
tag="${1:-r5h}"; out=$PWD/gpurun_out; mkdir -p $out
run() { sched="$1"; name="$2"; shift 2
  B200_SCHED=$sched timeout -s KILL 200 python bench.py "$@" --no-cpu-baseline --no-e2e --suite off --comparators off --steps 128 --warmup 16 \
    > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
for sc in 0,0 8,32 8,16 12,48 12,24 16,64 16,32 6,24; do
  n=${sc/,/_}
  run $sc ps64_$n --workload C3B --page-size 64
done
for sc in 0,0 8,32 12,48 16,64; do
  n=${sc/,/_}
  run $sc contig_$n --workload C3B --index-mode contiguous
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1]
for f in sorted(glob.glob(f"gpurun_out/{tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{f.split('/')[-1][len(tag)+1:-5]:18s} step {d['ms_per_step']*1e3:7.1f} us  kernel {d['roofline']['kernel_us_avg']:7.1f} us  "
              f"frac8T {d['roofline']['frac_of_8tbs']:.3f}  parity {d['parity_check']['ok'] if d.get('parity_check') else None}")
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 $out/${tag}_err.log 2>/dev/null
