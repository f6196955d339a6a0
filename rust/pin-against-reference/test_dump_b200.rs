// test_dump_b200.rs — dumps what pins rendiation_b200's oracle (and through it the CUDA kernels) against the REAL reference.
//
// Drop this file next to test.rs in shader/ray-tracing/src/backend/wavefront_compute/geometry/naive/ and add
//     #[cfg(test)] mod test_dump_b200;
// to naive/mod.rs (tools/pin_against_reference.sh does both), then
//     cargo test -p rendiation-device-ray-tracing dump_fixture_for_b200 -- --nocapture
// writes b200_fixture_dump.bin into the crate directory.  The file holds the INPUTS of the reference's own fixture
// (init_default_acceleration_structure, test.rs:9-225: every BLAS geometry as the builder received it, every TLAS instance)
// and, for each of the five TLASes x three ray-flag sets, what NaiveSahBvhCpu::traverse returned for a 64 x 64 pinhole grid:
// hit flag, geometry_idx, primitive_idx and the distance's bit pattern — plus the four visit counters of each pass.
// tools/compare_reference_dump.py rebuilds the scene from the dumped inputs (so the mesh generator's sin / cos need not agree
// between Rust and numpy), runs oracle/ over the same rays and compares record by record.
//
// Layout (little endian): magic "RDNDUMP1"; then sections  [name: 16 bytes, zero padded][count: u64][payload].
use std::io::Write;

use super::test::*;
use super::traverse_cpu::*;
use crate::backend::wavefront_compute::geometry::naive::*;

fn section(out: &mut Vec<u8>, name: &str, count: u64, payload: &[u8]) {
  let mut tag = [0u8; 16];
  tag[..name.len()].copy_from_slice(name.as_bytes());
  out.extend_from_slice(&tag);
  out.extend_from_slice(&count.to_le_bytes());
  out.extend_from_slice(payload);
}
fn f32s(v: impl IntoIterator<Item = f32>) -> Vec<u8> {
  v.into_iter().flat_map(|x| x.to_le_bytes()).collect()
}
fn u32s(v: impl IntoIterator<Item = u32>) -> Vec<u8> {
  v.into_iter().flat_map(|x| x.to_le_bytes()).collect()
}

#[test]
fn dump_fixture_for_b200() {
  const W: usize = 64;
  const H: usize = 64;
  const FAR: f32 = 100.;

  let (gpu, _) = futures::executor::block_on(GPU::new(Default::default())).unwrap();
  let system = NaiveSahBVHSystem::new(gpu);
  init_default_acceleration_structure(&system);
  // the reference binds TEST_TLAS_IDX only; the dump walks all five, so bind them all (handles are 0..5 in creation order)
  system.bind_tlas(&[TlasHandle(0), TlasHandle(1), TlasHandle(2), TlasHandle(3), TlasHandle(4)]);
  let _ = system.get_or_build_gpu_data();
  let inner = system.internal.read();
  let cpu_data = inner.cpu_data.as_ref().unwrap();

  let mut out: Vec<u8> = b"RDNDUMP1".to_vec();
  // ---- inputs
  for (b, blas) in inner.source.blas_data.iter().enumerate() {
    let geoms = blas.as_ref().expect("fixture deletes nothing");
    section(&mut out, "blas", geoms.len() as u64, &(b as u32).to_le_bytes());
    for g in geoms {
      match &g.geometry {
        BottomLevelAccelerationStructureBuildBuffer::Triangles { positions, indices } => {
          section(&mut out, "positions", positions.len() as u64, &f32s(positions.iter().flat_map(|p| [p.x, p.y, p.z])));
          let idx = indices.clone().unwrap_or_else(|| (0..positions.len() as u32).collect());
          section(&mut out, "indices", idx.len() as u64, &u32s(idx));
          section(&mut out, "geom_flags", 1, &(g.flags as u32).to_le_bytes());
        }
        BottomLevelAccelerationStructureBuildBuffer::AABBs { aabbs } => {
          section(&mut out, "aabbs", aabbs.len() as u64, &f32s(aabbs.iter().flat_map(|a| *a)));
          section(&mut out, "geom_flags", 1, &(g.flags as u32).to_le_bytes());
        }
      }
    }
  }
  for (t, tlas) in inner.source.tlas_data.iter().enumerate() {
    let insts = tlas.as_ref().expect("fixture deletes nothing");
    let mut payload = (t as u32).to_le_bytes().to_vec();
    for i in insts {
      let m: [f32; 16] = i.transform.into(); // column-major a1..d4 (math/algebra/src/mat/mat4.rs:9-14)
      payload.extend(f32s(m));
      payload.extend(u32s([
        i.instance_custom_index, i.mask, i.instance_shader_binding_table_record_offset, i.flags as u32, i.acceleration_structure_handle.0,
      ]));
    }
    section(&mut out, "tlas", insts.len() as u64, &payload);
  }

  // ---- rays (the reference's own recipe, test.rs:259-264) and results
  let mut dirs = Vec::with_capacity(W * H);
  for j in 0..H {
    for i in 0..W {
      let x = (i as f32 + 0.5) / W as f32 * 2. - 1.;
      let y = 1. - (j as f32 + 0.5) / H as f32 * 2.;
      dirs.push((vec3(x, y, -1.) - vec3(0., 0., 0.)).normalize());
    }
  }
  section(&mut out, "ray_dirs", dirs.len() as u64, &f32s(dirs.iter().flat_map(|d| [d.x, d.y, d.z])));

  let flag_sets: [(&str, u32); 3] = [
    ("cull_back", RayFlagConfigRaw::RAY_FLAG_CULL_BACK_FACING_TRIANGLES as u32),
    ("none", 0),
    ("first_hit", RayFlagConfigRaw::RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH as u32 | RayFlagConfigRaw::RAY_FLAG_CULL_BACK_FACING_TRIANGLES as u32),
  ];
  let counters = [&TRI_VISIT_COUNT, &TRI_HIT_COUNT, &BVH_VISIT_COUNT, &BVH_HIT_COUNT];
  for tlas_idx in 0..5u32 {
    for (name, flags) in flag_sets {
      for c in counters { c.store(0, std::sync::atomic::Ordering::Relaxed); }
      let mut payload = ShaderRayTraceCallStoragePayload::zeroed();
      payload.ray_flags = flags;
      payload.cull_mask = u32::MAX;
      payload.range = vec2(0., FAR);
      payload.tlas_idx = tlas_idx;
      payload.ray_origin = vec3(0., 0., 0.);
      let mut rec: Vec<u8> = Vec::with_capacity(dirs.len() * 16);
      for d in &dirs {
        payload.ray_direction = *d;
        match cpu_data.traverse(&payload, &mut |_| TEST_ANYHIT_BEHAVIOR) {
          Some(hit) => rec.extend(u32s([1, hit.geometry_idx, hit.primitive_idx, hit.distance.to_bits()])),
          None => rec.extend(u32s([0, u32::MAX, u32::MAX, FAR.to_bits()])),
        }
      }
      section(&mut out, &format!("t{tlas_idx}_{name}"), dirs.len() as u64, &rec);
      let c: Vec<u32> = counters.iter().map(|c| c.load(std::sync::atomic::Ordering::Relaxed)).collect();
      section(&mut out, &format!("c{tlas_idx}_{name}"), 4, &u32s(c)); // tri visit, tri hit, bvh visit, bvh hit
    }
  }
  std::fs::File::create("b200_fixture_dump.bin").unwrap().write_all(&out).unwrap();
  println!("wrote b200_fixture_dump.bin ({} bytes)", out.len());
}
