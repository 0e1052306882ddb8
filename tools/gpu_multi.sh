#!/usr/bin/env bash
# Multi-GPU check (gpurun --gpus N): the two-device single-process test, then the default bench under torchrun at
# every power of two up to the GPUs of the box (as the driver's scaling run does), both arms at N=2.
tag="${1:-r2m}"; out=$PWD/gpurun_out; mkdir -p $out
ngpu=$(python -c "import torch; print(torch.cuda.device_count())")
echo "gpus visible: $ngpu"
python -m pytest tests/test_abi_hardening_gpu.py -x -q -m gpu -k two_devices 2>&1 | tail -3
for n in 1 2 4 8; do
  [ $n -gt $ngpu ] && break
  if [ $n -eq 1 ]; then
    timeout -s KILL 600 python bench.py --gpus 1 --steps 20 --warmup 5 --comparators off > $out/${tag}_n1.json 2> $out/${tag}_n1.err
  else
    timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err
  fi
  echo "N=$n rc=$?"
done
python - "$tag" <<'PY'
import json, glob, sys
tag = sys.argv[1]
base = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{tag}_n{n}.json").read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        continue
    rows = [("headline " + d["config"]["workload"].split(":")[0], d["ms_per_step"] * 1e3, d["value"], d["e2e"]["value"] if d.get("e2e") else None, d["roofline"]["kernel_us_avg"])]
    for s in d.get("suite", []):
        if "error" in s: rows.append((s["workload"], None, None, None, None)); continue
        rows.append((s["workload"][:40], s["us_per_step"], s["value"], (s.get("e2e") or {}).get("value"), (s.get("roofline") or {}).get("kernel_us_avg")))
    for name, us, val, e2e, kus in rows:
        if us is None: print(f"N={n} {name}: ERROR"); continue
        if n == 1: base[name] = us
        eff = base.get(name, 0) / (n * us) if name in base else float("nan")
        print(f"N={n} {name:44s} step {us:8.1f} us  kernel {kus if kus else 0:8.1f}  value {val:12.0f}  e2e {e2e if e2e else 0:12.0f}  efficiency {eff:.3f}")
    if "allreduce_after_o_proj_us" in d: print(f"N={n} all-reduce after o_proj: {d['allreduce_after_o_proj_us']:.1f} us")
PY
for f in $out/${tag}_n*.err; do grep -v -i "warning\|warn(" $f | tail -2; done
