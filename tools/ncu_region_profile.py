"""Where a captured launch of k_trace_ordered_rounds spends its issue slots, by source region — no GPU needed.

    python tools/ncu_region_profile.py <report.ncu-rep> [--history] [--topup] [--share N] [--anyhit] [--json out.json]

Joins the SASS page of an `ncu --set full --import-source on` capture (instructions executed, thread instructions executed and
stall samples per SASS instruction) with the region of every instruction of the same instantiation in csrc/_obj/traverse.o
(tools/issue_model.py: the RDN_COST markers of traverse.cu / ordered_rounds.inc through nvdisasm's line info).  The object file
must be the build the capture ran.  Per region: share of the warp instructions, active lanes per instruction, share of the
stall samples."""
from __future__ import annotations

import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import issue_model  # noqa: E402


def instruction_regions(template_args: str, names: list[str]) -> list[int]:
    """region of every SASS instruction of the instantiation, in address order (the attribution of issue_model.static_counts)"""
    kernel_marks, tri_marks, (tri_first, tri_last) = issue_model.region_lines(names)
    tri_region = names.index("COST_TRI")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", issue_model.OBJ], cwd=tmp, check=True, capture_output=True)
        cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
        text = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True, check=True).stdout
    sections = re.split(r"\n//-+ \.text\.", text)
    sec = [s for s in sections if s.startswith("_ZN3rdn") and "k_trace_ordered_rounds" in s.split("\n")[0] and template_args in s.split("\n")[0]]
    if len(sec) != 1:
        raise RuntimeError(f"{len(sec)} instantiations match {template_args}")
    out, chain, chain_open = [], [], False
    for line in sec[0].split("\n"):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            if not chain_open:
                chain, chain_open = [], True
            for f, l in ((m.group(1), m.group(2)), (m.group(3), m.group(4))):
                if f and f.endswith("traverse.cu"):
                    chain.append(int(l))
                elif f and f.endswith("ordered_rounds.inc"):
                    chain.append(issue_model.INC_AT + int(l) * 1e-5)
            continue
        if not re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?[A-Z0-9_.]+", line):
            continue
        chain_open = False
        outer = chain[-1] if chain else kernel_marks[0][0]
        region = 0
        for mark_line, r in kernel_marks:
            if mark_line <= outer:
                region = r
        if region == tri_region:
            inner = next((l for l in chain if tri_first <= l <= tri_last), None)
            if inner is not None:
                for mark_line, r in tri_marks:
                    if mark_line <= inner:
                        region = r
        out.append(region)
    return out


def main():
    rep = sys.argv[1]
    share = int(sys.argv[sys.argv.index("--share") + 1]) if "--share" in sys.argv else 0
    names = issue_model.region_names()
    regions = instruction_regions(issue_model.mangled(3, inst_loop=True, share=share, anyhit="--anyhit" in sys.argv, history="--history" in sys.argv, topup="--topup" in sys.argv), names)
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
    lines = page.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    rows = [r for r in rows if r.get("Address", "").startswith("0x")]
    if len(rows) != len(regions):
        raise RuntimeError(f"the capture lists {len(rows)} SASS instructions, the object file {len(regions)}: not the same build")
    agg = {}
    for r, reg in zip(rows, regions):
        a = agg.setdefault(reg, [0, 0, 0])
        a[0] += int(r["Instructions Executed"]); a[1] += int(r["Thread Instructions Executed"]); a[2] += int(r["# Samples"] or 0)
    tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
    table = [{"region": names[reg], "warp_instructions": a[0], "share": a[0] / tot_i, "active_lanes": a[1] / a[0] if a[0] else 0.0,
              "stall_sample_share": a[2] / tot_s if tot_s else 0.0} for reg, a in sorted(agg.items(), key=lambda kv: -kv[1][0])]
    print(f"{'region':24s} {'warp instr':>12s} {'share':>7s} {'lanes':>6s} {'samples':>8s}")
    for t in table:
        print(f"{t['region']:24s} {t['warp_instructions']:12d} {100 * t['share']:6.1f}% {t['active_lanes']:6.1f} {100 * t['stall_sample_share']:7.1f}%")
    print(f"{'total':24s} {tot_i:12d}          {sum(a[1] for a in agg.values()) / tot_i:6.1f}")
    if "--json" in sys.argv:
        json.dump({"report": os.path.basename(rep), "regions": table, "warp_instructions": tot_i}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
