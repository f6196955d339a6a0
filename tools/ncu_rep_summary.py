"""Summarise one ncu --set full capture (.ncu-rep) of a kernel into a small JSON under profiles/.
    python tools/ncu_rep_summary.py <file.ncu-rep> <out.json> "<command / workload note>" """
import csv
import json
import subprocess
import sys

rep, out_path, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
out = {"note": note, "report": rep.split("/")[-1], "launches": []}
for r in data:
    m = {"kernel": r[hdr.index("Kernel Name")]}
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            try:
                m[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
            except ValueError:
                pass
    st = {}
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                st[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(r[i])
            except ValueError:
                pass
    m["warp_stall_cycles_per_issued_instruction"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
    m["dram_traffic_bytes"] = sum(m[k]["value"] * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
    if "lts__t_sectors.sum" in m:
        m["l2_traffic_bytes"] = m["lts__t_sectors.sum"]["value"] * 32
    ms = m["gpu__time_duration.sum"]["value"] * {"usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "ns": 1e-6, "second": 1e3, "s": 1e3}[m["gpu__time_duration.sum"]["unit"]]
    m["duration_ms"] = ms
    m["dram_gbs"] = m["dram_traffic_bytes"] / ms / 1e6
    if "l2_traffic_bytes" in m:
        m["l2_gbs"] = m["l2_traffic_bytes"] / ms / 1e6
    out["launches"].append(m)
json.dump(out, open(out_path, "w"), indent=1)
for m in out["launches"]:
    g = lambda k: m.get(k, {}).get("value")
    print(m["kernel"][:60], f"{m['duration_ms']:.4f} ms dram {m['dram_traffic_bytes']/1e6:.1f} MB ({m['dram_gbs']:.0f} GB/s) l2 {m.get('l2_traffic_bytes',0)/1e6:.1f} MB ({m.get('l2_gbs',0):.0f} GB/s) "
          f"inst {g('smsp__inst_executed.sum')} lanes {g('smsp__thread_inst_executed_per_inst_executed.ratio')} issue% {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} "
          f"regs {g('launch__registers_per_thread')} l1hit {g('l1tex__t_sector_hit_rate.pct')} local ld/st sectors {g('l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum')}/{g('l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum')}")
    print("   stalls:", m["warp_stall_cycles_per_issued_instruction"])
