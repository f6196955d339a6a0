"""The CUDA kernels executed on the CPU, thread for thread (tests/simt): the `-m gpu` test-suite run against
``tests/simt/_build/librdn_rt_emu.so`` — the same .cu sources compiled by g++ with a SIMT emulation header (one fiber per
CUDA thread, warp collectives, atomics, a host-memory stand-in for the CUDA runtime) behind the same C-ABI.

This checks kernel LOGIC without a GPU (traversal order, tie queue, refill votes, scans, the device builder's level loop); it
is test infrastructure, not a fallback: the package cannot load that library, only tests/conftest.py does under RDN_SIMT_EMU=1.
The GPU run of the same tests stays the parity gate.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(os.environ.get("RDN_SIMT_EMU") == "1", reason="already inside the emulated run")

# not meaningful or too slow on the emulator: stream overlap (58 launches of a million rays), a copy made with cuda-python,
# and the full-size BASELINE configurations (they pass — about a minute — and are run with RDN_SIMT_FULL=1)
SKIP_ALWAYS = ["back_to_back", "blob_adoption"]
SKIP_BIG = ["c2_c3_full_size", "million_triangles"]


def _run_emulated(selection, k_expr, env_extra=None, timeout=1500):
    env = dict(os.environ, RDN_SIMT_EMU="1", **(env_extra or {}))
    cmd = [sys.executable, "-m", "pytest", *selection, "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k", k_expr]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    summary = [line for line in r.stdout.splitlines() if " passed" in line]
    assert summary, tail
    return summary[-1]


def _not(names):
    return " and ".join(f"not {n}" for n in names)


def test_gpu_suite_passes_on_the_emulated_kernels():
    skip = SKIP_ALWAYS + ([] if os.environ.get("RDN_SIMT_FULL") == "1" else SKIP_BIG)
    summary = _run_emulated(["tests"], _not(skip))
    passed = int(summary.split(" passed")[0].split()[-1])
    assert passed >= 66, summary


def test_leaf_chains_walked_by_the_ordered_kernel():
    """Leaves of more than REF_LEAF_MAX_COUNT slots are stored as chains of wide nodes; with the real limit of 16 only
    geometry at the builder's depth limit produces one (and such geometry is irregular, so its rays take the reference-order
    walk).  Built with the limit lowered to 1, every two-triangle leaf and every multi-instance TLAS leaf is a chain, and
    the parity and fuzz tests must still hold bit for bit."""
    summary = _run_emulated(["tests/test_gpu_parity.py", "tests/test_gpu_fuzz.py"], _not(SKIP_ALWAYS + SKIP_BIG + ["hostile", "c4_instanced_full_size"]),
                            {"RDN_SIMT_DEFINES": "RDN_REF_LEAF_MAX_COUNT=1"})
    assert int(summary.split(" passed")[0].split()[-1]) >= 30, summary


def test_two_emulated_devices_shard_a_host_batch():
    summary = _run_emulated(["tests/test_gpu_parity.py"], "multi_device", {"RDN_SIMT_DEVICES": "2"})
    assert "1 passed" in summary, summary


def test_launch_rewriter():
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    import build_emu
    text, n = build_emu.rewrite_launches(
        "  k_a<true><<<static_cast<unsigned>(blocks), block, 0, stream>>>(scene, f(a, b), n);\n"
        "  if (n) k_b<<<dim3(gx, gy), 256>>>(p,\n      q);\n")
    assert n == 2 and "<<<" not in text
    assert "::simt::launch(dim3(static_cast<unsigned>(blocks)), dim3(block), [&]() { k_a<true>(scene, f(a, b), n); });" in text
    assert "::simt::launch(dim3(dim3(gx, gy)), dim3(256), [&]() { k_b(p, q); });" in text
    assert text.count("\n") == 3  # line numbers preserved


def test_issue_model_attributes_every_sass_instruction():
    """tools/issue_model.py, static half: every instruction of the default instantiation lands in a region, and the node step and
    the triangle iteration have plausible sizes (a change of the markers or of nvdisasm's output format shows up here)."""
    import shutil
    if not (shutil.which("nvdisasm") and shutil.which("cuobjdump")):
        pytest.skip("CUDA binary utilities not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import issue_model
    names = issue_model.region_names()
    static, total = issue_model.static_counts(issue_model.mangled(3, inst_loop=True), names)
    by = dict(zip(names, static))
    tri_family = ("COST_TRI", "COST_TRI_RANGE", "COST_TRI_U", "COST_TRI_V", "COST_TRI_HIT")
    copies = 3  # (the division below undoes the per-copy scaling only approximately: allow slack)
    assert abs(sum(static) + sum(by[n] for n in tri_family) * (copies - 1) - total) < 0.02 * total, (sum(static), total)
    assert 60 <= by["COST_NODE"] <= 110, by["COST_NODE"]
    assert 80 <= sum(by[n] for n in tri_family) <= 160, [by[n] for n in tri_family]


def test_two_gloo_ranks_trace_their_tiles_on_emulated_kernels():
    """the N > 1 launch path end to end on CPU: tile sharding, per-tile 2-D launches through the C ABI, gather on rank 0"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt", "two_rank_emulated.py")], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "identical to the oracle" in r.stdout, r.stdout[-1000:]


def _selftest(seed=None):
    code = ("import ctypes, sys; sys.path.insert(0, %r); import build_emu; L = ctypes.CDLL(build_emu.build()); "
            "L.simt_selftest_race.restype = ctypes.c_uint; print(L.simt_selftest_collectives(), L.simt_selftest_race(), L.simt_selftest_aggregate())"
            % os.path.join(ROOT, "tests", "simt"))
    env = dict(os.environ)
    env.pop("RDN_SIMT_SEED", None)
    if seed is not None:
        env["RDN_SIMT_SEED"] = str(seed)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    collectives, race, aggregate = r.stdout.split()[-3:]
    return int(collectives), int(race), int(aggregate)


def test_emulator_primitives_and_randomised_scheduling():
    """known answers for votes / shuffles / diverged masks / barriers / exited lanes, and the randomised scheduler really
    interleaves: a deliberately racy read-modify-write keeps all 64 increments under round-robin scheduling (every thread runs
    to completion) and loses some under RDN_SIMT_SEED"""
    ok, race, aggregate = _selftest()
    assert ok == 0 and race == 64, (ok, race)
    assert aggregate > 32, aggregate   # __activemask() groups the lanes that arrive together: the aggregated append ran with real groups
    lost = []
    for seed in (1, 2, 3):
        ok, race, aggregate = _selftest(seed)
        assert ok == 0 and aggregate >= 0, (seed, ok, aggregate)
        lost.append(64 - race)
    assert any(l > 0 for l in lost), lost


def test_parity_holds_under_randomised_scheduling():
    """the tie queue, the refill votes, the compaction's look-back and the fuzz scenes under a randomised interleaving of the
    lanes and warps (RDN_SIMT_SEED): the records must not depend on who runs when"""
    sel = "fixture or ties or flags or degenerate or instanced_scene or compaction or (fuzz and regular) or (fuzz and mixed)"
    for seed in (11, 12):
        summary = _run_emulated(["tests/test_gpu_parity.py", "tests/test_gpu_fuzz.py"], sel, {"RDN_SIMT_SEED": str(seed)})
        assert int(summary.split(" passed")[0].split()[-1]) >= 30, summary


def test_traversal_stack_overflow_is_reported_not_swallowed():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt", "stack_overflow_case.py")], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "stack overflow is reported" in r.stdout
