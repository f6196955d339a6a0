"""Instruction mix and stall attribution of the ordered kernel from an `ncu --set full --import-source on` report (SASS page):
per opcode class, the share of executed warp instructions, the average active lanes and the share of warp-stall samples; plus
the 20 instructions that collect the most samples.   python tools/ncu_sass_mix.py <report.ncu-rep> <out.json>"""
import collections, csv, json, re, subprocess, sys

rep, out_path = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
# the report holds several launches: take the first kernel block
blocks, cur = [], []
for line in raw:
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur: blocks.append(cur)
rows = list(csv.reader(blocks[0][1:]))
hdr, data = rows[0], rows[1:]
ci = {n: hdr.index(n) for n in ("Source", "Warp Stall Sampling (All Samples)", "Instructions Executed", "Thread Instructions Executed")}
CLASSES = [("fp32 add/mul", r"^(FADD|FMUL|FFMA)"), ("fp32 min/max", r"^FMNMX"), ("fp32 compare / select", r"^(FSETP|FSEL|FSET)"),
           ("fp32 divide / sqrt (MUFU + fix-up)", r"^(MUFU|FCHK)"), ("global load", r"^LDG"), ("global store / atomic", r"^(STG|ATOMG|RED|ATOM)"),
           ("local memory (stack / spills)", r"^(LDL|STL)"), ("constant / param load", r"^(LDC|ULDC|LDCU)"),
           ("integer / logic / move", r"^(IMAD|IADD|LEA|LOP|SHF|MOV|SEL|ISETP|PLOP|VIADD|IABS|POPC|FLO|PRMT|I2F|F2I|CS2R|S2R|S2UR|R2UR|UMOV|ULOP|UIADD|USHF|UISETP|UFLO|UPOPC|UIMAD|ULEA|USEL|UPLOP|R2P|P2R)"),
           ("branch / convergence", r"^(BRA|BSSY|BSYNC|WARPSYNC|EXIT|CALL|RET|YIELD|NOP|BREAK|BMOV|BAR|BPT|JMP|BRX)"), ("vote / shuffle", r"^(VOTE|VOTEU|SHFL|MATCH|REDUX)")]
agg = collections.OrderedDict((n, [0, 0, 0]) for n, _ in CLASSES)
agg["other"] = [0, 0, 0]
tot = [0, 0, 0]
top = []
for r in data:
    src = r[ci["Source"]].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", src)
    inst, thr, smp = (int(float(r[ci[k]] or 0)) for k in ("Instructions Executed", "Thread Instructions Executed", "Warp Stall Sampling (All Samples)"))
    name = next((n for n, pat in CLASSES if re.match(pat, op)), "other")
    for a in (agg[name], tot):
        a[0] += inst; a[1] += thr; a[2] += smp
    top.append((smp, inst, thr, src))
top.sort(reverse=True)
res = {"kernel": blocks[0][0].split('","')[1].rstrip('",'), "warp_instructions": tot[0], "avg_active_lanes": tot[1] / max(tot[0], 1), "stall_samples": tot[2],
       "classes": [{"class": n, "share_of_warp_instructions": a[0] / tot[0], "avg_active_lanes": a[1] / max(a[0], 1), "share_of_stall_samples": a[2] / max(tot[2], 1)}
                   for n, a in agg.items() if a[0]],
       "top_by_stall_samples": [{"sass": s, "share_of_stall_samples": smp / max(tot[2], 1), "executed": inst, "avg_active_lanes": thr / max(inst, 1)} for smp, inst, thr, s in top[:20]]}
json.dump(res, open(out_path, "w"), indent=1)
for c in res["classes"]:
    print(f"{c['class']:40s} {100*c['share_of_warp_instructions']:5.1f} % of instructions  {c['avg_active_lanes']:4.1f} lanes  {100*c['share_of_stall_samples']:5.1f} % of stall samples")
for t in res["top_by_stall_samples"][:12]:
    print(f"{100*t['share_of_stall_samples']:5.2f} %  {t['avg_active_lanes']:4.1f} lanes  {t['sass'][:90]}")
