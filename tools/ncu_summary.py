#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (key metrics per profiled launch + top stall sites of the
first launch) or of an ncu launch-list CSV (--launches).  Output goes under profiles/.
  tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
  tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/launches.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg, order = None, collections.defaultdict(list), []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        name = d["Kernel Name"].split("(")[0][-70:]
        agg[name].append(v)
        order.append((name, v))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {len(order)} launches, {tot:.1f} us total (ncu per-launch times: cold-cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} n={len(v):4d} avg={sum(v) / len(v):9.2f} us  share={sum(v) / tot * 100:5.1f}%")
    print("# first 12 launches in order:")
    for name, v in order[:12]:
        print(f"   {v:9.2f} us  {name}")


def report(path):
    raw = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for n, r in enumerate(raw[2:]):
        d = dict(zip(hdr, r))
        print(f"== launch {n}: {d['Kernel Name'][:110]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] not in ("", "0"):
                print(f"   {k:78s} {d[k]:>16s} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv", "--kernel-id", ":::1"]))))
    hi = next((i for i, r in enumerate(src) if r and r[0] == "Address"), None)
    if hi is None:
        return
    h = src[hi]
    ix = {k: h.index(k) for k in ("Source", "# Samples", "stall_long_sb", "stall_barrier", "stall_short_sb",
                                  "stall_math", "stall_mio", "stall_wait", "Instructions Executed")}
    data = []
    for n, r in enumerate(src[hi + 1:]):
        try:
            data.append((int(r[ix["# Samples"]] or 0), n, r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"== top stall sites of launch 0 ({tot} samples, {len(data)} SASS instructions)")
    for s, n, r in sorted(data, key=lambda d: -d[0])[:16]:
        print(f"   {s:6d} {s / tot * 100:5.1f}%  long={r[ix['stall_long_sb']]:>5s} bar={r[ix['stall_barrier']]:>5s} "
              f"ssb={r[ix['stall_short_sb']]:>5s} math={r[ix['stall_math']]:>5s} mio={r[ix['stall_mio']]:>5s} "
              f"wait={r[ix['stall_wait']]:>5s}  {r[ix['Source']].strip()[:70]}")
    ops = collections.Counter()
    for s, n, r in data:
        op = r[ix["Source"]].strip().split()
        if op:
            name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            ops[name.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    print("== executed warp-instructions by opcode (top 12):", ", ".join(f"{k}={v}" for k, v in ops.most_common(12)))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1])
