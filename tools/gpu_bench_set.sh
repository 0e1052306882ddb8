#!/usr/bin/env bash
# Development sweep run on the B200 box through gpurun: bench lines for the BASELINE.json decode
# configs (per-GPU shard shapes emulated with --shard) into gpurun_out/<tag>_*.json.
tag="${1:-dev}"
out=gpurun_out
mkdir -p $out
run() { name="$1"; shift; python bench.py "$@" --no-cpu-baseline > $out/${tag}_${name}.json 2>> $out/${tag}_err.log || echo "FAILED $name" >> $out/${tag}_err.log; }
run c2 --workload C2
run c2_contig --workload C2 --index-mode contiguous
run c3b --workload C3B
run c3u --workload C3U
run c3 --workload C3
run c5_tp8 --workload C5 --shard 8 --steps 512
run c2_tp8 --workload C2 --shard 8
run c3u_tp8 --workload C3U --shard 8
run c3b_tp8 --workload C3B --shard 8
run c3u_tp2 --workload C3U --shard 2
run c3u_tp4 --workload C3U --shard 4
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/${tag}_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j["roofline"]
    print(f"{f.split('/')[-1]:28s} step {j['ms_per_step']*1e3:8.1f}us  {j['value']/1e6:7.3f} Mtok/s  step-GB/s {j['hbm_gbs_step']:7.0f}  "
          f"kernel {r['kernel_us_avg']:8.1f}us {r['achieved']:7.0f} GB/s ({r['frac']*100:5.1f}% meas, {r['frac_of_8tbs']*100:5.1f}% of 8TB/s)  "
          f"e2e {j['e2e']['value']/1e6 if j['e2e'] else 0:6.3f} Mtok/s  clk {j['clocks']['sm_mhz']} {j['clocks']['reasons']}")
PY
tail -5 $out/${tag}_err.log 2>/dev/null
