"""bench.py's reference arm (the reference's CPU traversal = the oracle port, timed on the host cores) runs here without a GPU:
it must print exactly one JSON line on stdout carrying the keys the driver reads, and refuse politely when asked for our arm
without a CUDA device (the product has no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "closest-hit Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "BASELINE configs[1]" in d["config"]["workload"] and d["dtype"] == "f32" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_our_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_both_arms_describe_the_workload_with_the_same_config():
    """the driver compares the `config` objects of the two arms: they come from one function"""
    sys.path.insert(0, ROOT)
    import bench
    c = bench.base_config()
    assert set(c) == {"workload", "rays_per_gpu_per_step", "triangles", "ray_flags"} and c["rays_per_gpu_per_step"] == 1920 * 1080
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": base_config()') == 2  # reference arm and our arm, configs[1]


def test_the_kernel_named_in_the_roofline_is_the_one_the_launcher_selects_for_grids():
    """bench.py quotes ncu numbers of SHIPPED_ORDERED_KERNEL; it must be the instantiation launch_trace_ordered picks for device-resident
    grid launches without an any-hit stage — the one that keeps a tile history (template arguments as ncu prints them: every
    parameter, booleans as 0 / 1)"""
    import re
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "rendiation_b200", "csrc", "traverse.cu")).read()
    m = re.search(r"// a grid with a tile history \(capi\.cu\)\n\s*fn = k_trace_ordered_rounds<([^>]*)>;", src)
    macros = {name: re.search(r"#define %s (\d+)" % name, src).group(1) for name in ("RDN_ORDERED_K", "RDN_ORDERED_MINB")}
    args = [macros.get(a.strip(), a.strip()) for a in m.group(1).split(",")]
    n_params = len(re.search(r"template <(int K, int MINB,[^>]*)>\n__global__", src).group(1).split(","))
    args += ["false"] * (n_params - len(args))   # defaulted trailing parameters
    as_ncu = ", ".join({"true": "1", "false": "0", "SHARE_NEVER": "0", "SHARE_ALWAYS": "1", "SHARE_LATE": "2"}.get(a, a) for a in args)
    assert bench.SHIPPED_ORDERED_KERNEL == f"k_trace_ordered_rounds<{as_ncu}>", (bench.SHIPPED_ORDERED_KERNEL, as_ncu)


def test_reference_arm_of_config_5_runs_a_bounded_sample():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c5", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["impl"] == "reference" and "configs[4]" in d["config"]["workload"] and d["scaling"] == "strong" and d["value"] > 0
