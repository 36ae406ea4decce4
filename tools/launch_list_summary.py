#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean duration and share of the
serialised kernel time per kernel (the ffma_bench_kernel peak microbenchmark and the sim_kernel frame generator run outside the timed region and are left out).

    python tools/launch_list_summary.py gpurun_out/r01f_launches.csv "<command that was profiled>" > profiles/r01f_launch_list_summary.json
"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    command = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik]).strip()
        if "ffma_bench_kernel" in name or name.startswith("sim_kernel"):
            continue  # peak microbenchmark / device-side frame generation: both run before the timed region
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    kernels = [{"kernel": k, "launches": cnt[k], "mean_us": round(tot[k] / cnt[k], 3), "share_pct": round(100 * tot[k] / total, 2)}
               for k in sorted(tot, key=lambda k: -tot[k])]
    print(json.dumps({"command": command,
                      "note": "ffma_bench_kernel (peak microbenchmark) and sim_kernel (device-side frame generation), both outside the timed region, excluded from shares; per-launch "
                              "times are cold-cache and serialised; in the live step the other kernels overlap the scorer from other streams",
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
