"""The Rust side of the boundary (rust/): a `-sys` crate whose declarations are generated from include/rdn_rt.h, a safe wrapper with
the reference's method names, and the kit that pins the oracle against the real reference.  No Rust toolchain exists in the build
image, so these tests hold what can be held without one: the generated file is current, every exported symbol is declared, the
crate builds the same sources with the same flags as the Python build, and the dump format of the pinning kit round-trips."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_sys_crate_declarations_are_generated_from_the_header():
    import gen_rust_sys
    committed = open(gen_rust_sys.OUT).read()
    assert committed == gen_rust_sys.generate(), "include/rdn_rt.h changed: run python tools/gen_rust_sys.py"


def test_every_exported_symbol_is_declared_in_the_sys_crate():
    from rendiation_b200 import api
    src = open(os.path.join(ROOT, "rust", "rendiation-rt-b200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (rdn_\w+)\(", src))
    assert set(api.EXPORTED_SYMBOLS) <= declared, sorted(set(api.EXPORTED_SYMBOLS) - declared)
    header = open(os.path.join(ROOT, "include", "rdn_rt.h")).read()
    in_header = set(re.findall(r"^(?:int|void|uint32_t|const char \*)\s*(rdn_\w+)\(", header, flags=re.M))
    assert declared == in_header, (sorted(declared - in_header), sorted(in_header - declared))


def test_record_layouts_match_the_ctypes_mirror():
    """sizes of the #[repr(C)] records as Rust lays them out (fields in order, natural alignment) == the ctypes structures the tests use"""
    import ctypes as C

    from rendiation_b200 import api
    src = open(os.path.join(ROOT, "rust", "rendiation-rt-b200-sys", "src", "lib.rs")).read()
    size = {"f32": 4, "u32": 4, "i32": 4, "u64": 8, "f64": 8, "u8": 1}

    def rust_sizeof(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, flags=re.S).group(1)
        off, align = 0, 1
        for ty in re.findall(r"pub \w+: ([^,]+),", body):
            m = re.match(r"\[(\w+); (\d+)\]", ty)
            base, count = (m.group(1), int(m.group(2))) if m else (ty, 1)
            a = 8 if base.startswith("*") else size[base]
            off = (off + a - 1) // a * a + a * count
            align = max(align, a)
        return (off + align - 1) // align * align

    for rust_name, mirror in (("rdn_launch", api._Launch), ("rdn_trace_stats", api._TraceStats), ("rdn_counters", api._Counters), ("rdn_pinhole", api._Pinhole),
                              ("rdn_camera", api._Camera), ("rdn_bounce", api._Bounce), ("rdn_kernel_times", api._KernelTimes)):
        assert rust_sizeof(rust_name) == C.sizeof(mirror), rust_name
    assert rust_sizeof("rdn_ray") == api.RAY_DTYPE.itemsize == 32 and rust_sizeof("rdn_hit") == api.HIT_DTYPE.itemsize == 32
    assert rust_sizeof("rdn_instance") == 84


def test_build_rs_compiles_what_the_python_build_compiles():
    from rendiation_b200 import build
    rs = open(os.path.join(ROOT, "rust", "rendiation-rt-b200-sys", "build.rs")).read()
    sources = re.findall(r'"([\w]+\.(?:cu|cpp))"', rs.split("const SOURCES")[1].split("];")[0])
    assert sources == build.SOURCES
    for flag in ("arch=compute_100a,code=sm_100a", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-ffp-contract=off"):
        assert flag in rs and any(flag in f for f in build.NVCC_FLAGS), flag


def test_safe_wrapper_keeps_the_reference_method_names():
    src = open(os.path.join(ROOT, "rust", "rendiation-rt-b200", "src", "lib.rs")).read()
    for name in ("create_bottom_level_acceleration_structure", "delete_bottom_level_acceleration_structure", "create_top_level_acceleration_structure",
                 "delete_top_level_acceleration_structure", "bind_tlas", "bind_tlas_max_len", "trace_closest_batch"):
        assert re.search(r"pub fn %s\(" % name, src), name
    prov = open(os.path.join(ROOT, "rust", "rendiation-rt-b200", "src", "provider.rs")).read()
    assert "impl GPUAccelerationStructureSystemProvider for B200BvhSystem" in prov


def test_pinning_kit_dump_format_round_trips():
    """tools/compare_reference_dump.py --self-test: the oracle writes the reference test's dump format and the comparison reads it back"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "compare_reference_dump.py"), "--self-test"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PINNED" in r.stdout, (r.stdout + r.stderr)[-2000:]
    assert r.stdout.count("records identical") == 15
    rs = open(os.path.join(ROOT, "rust", "pin-against-reference", "test_dump_b200.rs")).read()
    for section in ("positions", "indices", "geom_flags", "tlas", "ray_dirs", "RDNDUMP1"):
        assert section in rs
