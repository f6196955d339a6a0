#!/bin/bash
# Pins this repository's oracle (and, through the GPU suite, the CUDA kernels) against the REAL reference.  Needs what this image
# lacks: the reference's Rust toolchain (rust-toolchain.toml: nightly-2026-08-08) and network access for its crates.
#
#   tools/pin_against_reference.sh /path/to/rendiation [/path/to/this/repo]
#
# 1. adds rust/pin-against-reference/test_dump_b200.rs to the reference's naive geometry backend as a test module (two lines of patch,
#    reverted afterwards), 2. runs the reference's own test_cpu_triangle and the dump test, 3. byte-compares trace_cpu.pbm and the
#    printed visit counters with tests/golden/, 4. rebuilds the fixture from the dumped inputs in oracle/ and compares every record
#    of all five TLASes x three flag sets (ids exact, distance bit patterns).
set -euo pipefail
REF=${1:?path to a checkout of mikialex/rendiation}
HERE=${2:-$(cd "$(dirname "$0")/.." && pwd)}
NAIVE=$REF/shader/ray-tracing/src/backend/wavefront_compute/geometry/naive
CRATE=$REF/shader/ray-tracing
cp "$HERE/rust/pin-against-reference/test_dump_b200.rs" "$NAIVE/test_dump_b200.rs"
cp "$NAIVE/mod.rs" "$NAIVE/mod.rs.b200bak"
trap 'mv "$NAIVE/mod.rs.b200bak" "$NAIVE/mod.rs"; rm -f "$NAIVE/test_dump_b200.rs"' EXIT
# test.rs items are pub(crate) already; the module just has to exist next to `mod test;`
sed -i '0,/^mod test;/s//mod test;\n#[cfg(test)]\nmod test_dump_b200;/' "$NAIVE/mod.rs"
( cd "$REF" && cargo test -p rendiation-device-ray-tracing test_cpu_triangle -- --nocapture --test-threads 1 | tee "$CRATE/test_cpu_triangle.log" )
( cd "$REF" && cargo test -p rendiation-device-ray-tracing dump_fixture_for_b200 -- --nocapture --test-threads 1 )
python "$HERE/tools/compare_reference_dump.py" --pbm "$CRATE/trace_cpu.pbm" --log "$CRATE/test_cpu_triangle.log" --dump "$CRATE/b200_fixture_dump.bin"
