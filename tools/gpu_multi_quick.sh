#!/usr/bin/env bash
# Quick scaling check of the headline only: N = 1 and N = all GPUs of the box, driver flags (--steps 20 --warmup 5),
# plus a longer run (--steps 256) at N = all to separate per-graph overheads from the steady state
tag="${1:-r2q}"; out=$PWD/gpurun_out; mkdir -p $out
ngpu=$(python -c "import torch; print(torch.cuda.device_count())")
F="--suite off --comparators off --no-cpu-baseline --no-parity"
timeout -s KILL 300 python bench.py --gpus 1 --steps 20 --warmup 5 $F > $out/${tag}_n1.json 2> $out/${tag}_n1.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $ngpu --steps 20 --warmup 5 $F > $out/${tag}_n${ngpu}.json 2> $out/${tag}_n${ngpu}.err
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $ngpu --steps 256 --warmup 16 $F --no-e2e > $out/${tag}_n${ngpu}_long.json 2> $out/${tag}_n${ngpu}_long.err
python - "$tag" "$ngpu" <<'PY'
import json, sys
tag, n = sys.argv[1], int(sys.argv[2])
rd = lambda k: json.loads([l for l in open(f"gpurun_out/{tag}_n{k}.json").read().strip().splitlines() if l.startswith("{")][-1])
a = rd(1)
for k in (str(n), f"{n}_long"):
    b = rd(k)
    print(f"N=1 step {a['ms_per_step']*1e3:.1f} us kernel {a['roofline']['kernel_us_avg']:.1f}; N={k} steps {b['steps']} step {b['ms_per_step']*1e3:.1f} us kernel(rank0) {b['roofline']['kernel_us_avg']:.1f}; efficiency {a['ms_per_step']/(n*b['ms_per_step']):.3f}")
    print("   per-rank step us:", [round(x, 1) for x in b.get("step_us_per_rank", [])], "clocks", b["clocks"].get("sm_mhz"), b["clocks"].get("reasons"))
PY
