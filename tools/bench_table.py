#!/usr/bin/env python
"""One line per bench JSON file: step time, tokens/s, kernel GB/s and roofline fractions."""
import json
import sys

for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable", e)
        continue
    r = j["roofline"]
    if r["bound"] != "hbm":
        print(f"{f.split('/')[-1]:30s} step {j['ms_per_step'] * 1e3:9.1f}us  {r['achieved']:7.1f} TFLOP/s ({r['frac'] * 100:5.1f}% of measured "
              f"sustained, {r.get('frac_of_burst', r['frac']) * 100:5.1f}% of burst)")
        continue
    e2e = j["e2e"]["value"] / 1e6 if j.get("e2e") else 0
    print(f"{f.split('/')[-1]:30s} step {j['ms_per_step'] * 1e3:8.1f}us {j['value'] / 1e6:7.3f} Mtok/s step-GB/s {j['hbm_gbs_step']:6.0f} | "
          f"kernel {r['kernel_us_avg']:7.1f}us {r['achieved']:6.0f} GB/s ({r['frac'] * 100:5.1f}% meas {r['frac_of_8tbs'] * 100:5.1f}% 8TB/s) | "
          f"e2e {e2e:6.3f} Mtok/s | clk {j['clocks']['sm_mhz']} {j['clocks']['reasons']}")
