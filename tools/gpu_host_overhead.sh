#!/usr/bin/env bash
# Host side of a layer-call: where the e2e step goes (copies, hook, replay, sync), and eager launches against graph
# replay at the shard shapes where a layer-call is shortest.
tag="${1:-r2h}"; out=$PWD/gpurun_out; mkdir -p $out
python tools/e2e_breakdown.py C3B 2>&1 | grep "us per step" | tee $out/${tag}_e2e_c3b.txt
python tools/e2e_breakdown.py C2 2>&1 | grep "us per step" | tee $out/${tag}_e2e_c2.txt
for w in C2 C3B; do for m in graph graph1 eager; do
  python bench.py --workload $w --shard 8 --mode $m --steps 256 --warmup 16 --suite off --comparators off --no-cpu-baseline --no-parity > $out/${tag}_${w}_tp8_$m.json 2>> $out/${tag}_err.log
done; done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*_tp8_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1][len(sys.argv[1])+1:-5], f"step {d['ms_per_step']*1e3:.1f} us  e2e {d['e2e']['ms_per_step']*1e3:.1f} us")
PY
