#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU) into the few numbers the roofline discussion needs.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/<name>.txt
"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pat and not pat.search(d["Kernel Name"]):
            continue
        print("kernel:", d["Kernel Name"])
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued")),
                        reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp stall samples:", ", ".join(f"{k.split('stalled_')[1]} {100 * v / tot:.0f}%" for v, k in stalls[:6]))
        print()
    src = page(rep, "source")
    body = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        body.append(r)
    top = sorted(range(len(body)), key=lambda i: -int(body[i][2]) if body[i][2].isdigit() else 0)[:10]
    tot = sum(int(r[2]) for r in body if r[2].isdigit()) or 1
    print("top stall sites of the first kernel (SASS, % of samples, previous instruction):")
    for i in top:
        print(f"  {100 * int(body[i][2]) / tot:5.1f}%  {body[i][1].strip()[:70]:70s} <- {body[i - 1][1].strip()[:60]}")


if __name__ == "__main__":
    main()
