"""CPU checks of bench.py's host logic: the reference arm prints one well-formed JSON line whose `config` is the very
dict the b200 arm prints for the same flags (the driver compares the two arms), and the workload bookkeeping agrees
with SURVEY 8d's algorithmic-bytes formula."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(*argv):
    import bench
    old = sys.argv
    sys.argv = ["bench.py", *argv]
    try:
        return bench.parse_args()
    finally:
        sys.argv = old


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same declarative config as the b200 arm builds from the same flags
    import bench
    from harness import synth
    args = _args("--workload", "C1", "--steps", "2", "--warmup", "1")
    w = synth.WORKLOADS["C1"]
    assert line["config"] == json.loads(json.dumps(bench.workload_config(w, 1, args, bench.static_info(args, w, 1))))


def test_default_workload_is_the_north_star_shape_and_bytes_follow_survey_8d():
    import bench
    from harness import synth
    args = _args()
    assert args.workload == "C3B" == bench.HEADLINE
    w = synth.WORKLOADS[args.workload]
    assert (w.hq_total, w.hkv_total, w.head_dim, w.batch, w.ctx, w.kv_dtype) == (32, 8, 128, 256, 4096, torch.bfloat16)
    info = bench.static_info(args, w, 1)
    tot = sum(info["seq_lens"])
    want = tot * 8 * (128 + 128) * 2 + 2 * 256 * 32 * 128 * 2 + tot * 4 + 256 * 16
    assert synth.decode_algorithmic_bytes(info["seq_lens"], 32, 8, 128, 128, 2) == want == 4303360000
    # sharding keeps the page table and divides the heads
    for tp in (2, 4, 8):
        i2 = bench.static_info(args, w, tp)
        assert i2["seq_lens"] == info["seq_lens"] and i2["hkv"] == 8 // tp and i2["hq"] == 32 // tp
    assert bench.static_info(args, w, 8)["fused_kv"] is False and bench.static_info(_args("--workload", "C3U"), synth.WORKLOADS["C3U"], 8)["fused_kv"]
