#!/usr/bin/env python
"""Per-source-line view of an ncu capture: dynamic warp-instructions, average live threads and stall samples per line.

ncu's SASS page (instructions executed / stall samples per instruction) is joined, instruction by instruction, with
`nvdisasm -g` of the same cubin extracted from libbeng.so, which carries the file:line of every instruction.  Runs here,
without a GPU.  Subroutines without line info (the IEEE division slow paths) inherit the last line seen: look for
CALL targets with --calls.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep traffic traffic_step_kernelILi9ELi3ELi72ELb0 [--per N] [--min M] [--calls]
      (report, .cu basename, mangled-name fragment; --per: divide counts by N, e.g. the grid size)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "custom_gymnasium_environments_b200", "libbeng.so")


def main():
    rep, unit, frag = sys.argv[1:4]
    opts = sys.argv[4:]
    per = float(opts[opts.index("--per") + 1]) if "--per" in opts else 1.0
    lo = float(opts[opts.index("--min") + 1]) if "--min" in opts else 0.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the page lists one block per kernel: "Kernel Name", header, instructions...
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    blk = blocks[0]
    h = blk["rows"][0]
    ie, at, ss = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    body = [r for r in blk["rows"][1:] if len(r) > ie and r[ie].isdigit()]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", unit, LIB], cwd=td, check=True, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\.\S*" + re.escape(frag), l))
    seq, line = [], None
    for l in dis[start + 1:]:
        if re.match(r"\s*\.section", l):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append((line, m.group(2)))
    if len(seq) != len(body):
        sys.exit(f"SASS length mismatch: report {len(body)} vs library {len(seq)} -- the capture is from another build")
    agg, thr, st = collections.Counter(), collections.Counter(), collections.Counter()
    for (ln, txt), r in zip(seq, body):
        agg[ln] += int(r[ie]); thr[ln] += int(r[at]); st[ln] += int(r[ss])
        if "--calls" in opts and "CALL" in txt and int(r[ie]):
            print(f"call {ln} x{int(r[ie]) / per:.2f} thr {int(r[at]) / int(r[ie]):.1f}: {txt[:90]}")
    srcs = {}
    tot, tst = sum(agg.values()), sum(st.values()) or 1
    print(f"kernel {blk['name'][:100]}\n{len(body)} SASS instructions, {tot / per:.1f} dynamic warp-instructions per unit")
    for ln in sorted(agg, key=lambda k: (k is None, k)):
        if agg[ln] / per < lo and 100 * st[ln] / tst < 1.0:
            continue
        text = ""
        if ln:
            if ln[0] not in srcs:
                cands = [os.path.join(dp, ln[0]) for dp, _, fs in os.walk(os.path.join(ROOT, "custom_gymnasium_environments_b200")) if ln[0] in fs]
                srcs[ln[0]] = open(cands[0]).read().split("\n") if cands else []
            if ln[1] - 1 < len(srcs[ln[0]]):
                text = srcs[ln[0]][ln[1] - 1].strip()[:96]
        tag = f"{ln[0]}:{ln[1]}" if ln else "?"
        print(f"{tag:>24s} {agg[ln] / per:9.1f} thr {thr[ln] / max(agg[ln], 1):4.1f} stall {100 * st[ln] / tst:4.1f}%  {text}")


if __name__ == "__main__":
    main()
