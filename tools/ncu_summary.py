"""Summarise a gpu_round.sh session (gpurun_out/<tag>/) into profiles/: per-launch list of the bench command and the
--set full capture of the top kernel.   python tools/ncu_summary.py <tag> <suffix>"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, suffix = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = os.path.join(ROOT, "gpurun_out", tag)
prof = os.path.join(ROOT, "profiles")
raw = subprocess.run(["ncu", "-i", os.path.join(d, "prof_ordered.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
out = {"source": open(os.path.join(d, "ncu_full.log")).read().strip().splitlines()[-3:], "command":
       "ncu --set full --clock-control none --import-source on -k regex:k_trace_ordered -s 3 -c 2 python tools/kbench.py c2 3 "
       "(B200; BASELINE configs[1]: 2,073,600 primary rays vs the 1,002,528-triangle torus; cold caches per replay)",
       "kernel": data[0][hdr.index("Kernel Name")], "launches": []}
for r in data:
    m = {}
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            m[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
    st = {}
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            st[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(r[i])
    m["warp_stall_cycles_per_issued_instruction"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
    m["dram_traffic_bytes"] = sum(m[k]["value"] * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    m["l2_traffic_bytes"] = m["lts__t_sectors.sum"]["value"] * 32
    out["launches"].append(m)
json.dump(out, open(os.path.join(prof, f"ncu_{suffix}_k_trace_ordered_c2.json"), "w"), indent=1)

rows = [r for r in csv.reader(open(os.path.join(d, "launches.csv"))) if len(r) > 5]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
L = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 1 "
                "(per-launch times are cold-cache and serialised: compare shares, not absolutes)",
     "kernels": [{"name": k, "launches": len(v), "mean_us": sum(v) / len(v) / 1e3, "min_us": min(v) / 1e3, "max_us": max(v) / 1e3,
                  "share_of_listed_gpu_time": sum(v) / tot} for k, v in agg.items()]}
json.dump(L, open(os.path.join(prof, f"launches_{suffix}_bench_c2.json"), "w"), indent=1)
for f, t in (("launches.csv", f"launches_{suffix}_bench_c2.csv"), ("bench.json", f"bench_{suffix}.json"), ("bench_reference.json", f"bench_{suffix}_reference_arm.json")):
    if os.path.exists(os.path.join(d, f)):
        open(os.path.join(prof, t), "w").write(open(os.path.join(d, f)).read())
m = out["launches"][-1]
print(json.dumps({k: (v["value"] if isinstance(v, dict) and "value" in v else v) for k, v in m.items()}, indent=1))
print(json.dumps(L["kernels"], indent=1))
