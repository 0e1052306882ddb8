#!/usr/bin/env bash
# Cascade on/off comparison on the shared-prefix workload (C3) and its overhead on unshared ones (C3U, C3B, C5 shard).
tag="${1:-dev}"; out=gpurun_out; mkdir -p $out
run() { name="$1"; shift; timeout -s KILL 120 python bench.py "$@" --no-cpu-baseline > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run c3_on --workload C3
run c3_off --workload C3 --cascade off
run c3u_on --workload C3U
run c3u_off --workload C3U --cascade off
run c3_tp8_on --workload C3 --shard 8
run c3_tp8_off --workload C3 --shard 8 --cascade off
run c5_tp8_on --workload C5 --shard 8 --steps 512
run c5_tp8_off --workload C5 --shard 8 --steps 512 --cascade off
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j["roofline"]
    print(f"{f.split('/')[-1]:28s} step {j['ms_per_step']*1e3:8.1f}us  {j['value']/1e6:7.3f} Mtok/s  attn legs {r['kernel_us_avg']:8.1f}us  "
          f"e2e {j['e2e']['value']/1e6 if j['e2e'] else 0:6.3f} Mtok/s  shared {j['config']['cascade']['shared_prefix_tokens_detected']}  launches {j['gpu_launches']}")
PY
tail -5 $out/${tag}_err.log 2>/dev/null
