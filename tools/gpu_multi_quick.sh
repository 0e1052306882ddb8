#!/usr/bin/env bash
# Quick scaling check of the headline only: N = 1 and N = all GPUs of the box, driver flags (--steps 20 --warmup 5)
tag="${1:-r2q}"; out=$PWD/gpurun_out; mkdir -p $out
ngpu=$(python -c "import torch; print(torch.cuda.device_count())")
F="--steps 20 --warmup 5 --suite off --comparators off --no-cpu-baseline"
timeout -s KILL 300 python bench.py --gpus 1 $F > $out/${tag}_n1.json 2> $out/${tag}_n1.err
for n in 4 $ngpu; do
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n $F > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err
done
python - "$tag" "$ngpu" <<'PY'
import json, sys
tag, n = sys.argv[1], int(sys.argv[2])
rd = lambda k: json.loads([l for l in open(f"gpurun_out/{tag}_n{k}.json").read().strip().splitlines() if l.startswith("{")][-1])
a = rd(1)
for k in (4, n):
    b = rd(k)
    print(f"N=1 step {a['ms_per_step']*1e3:.1f} us kernel {a['roofline']['kernel_us_avg']:.1f}; N={k} step {b['ms_per_step']*1e3:.1f} us kernel {b['roofline']['kernel_us_avg']:.1f}; efficiency {a['ms_per_step']/(k*b['ms_per_step']):.3f}; e2e {a['e2e']['value']:.0f} -> {b['e2e']['value']:.0f}")
PY
