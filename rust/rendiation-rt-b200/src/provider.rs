//! `impl GPUAccelerationStructureSystemProvider for B200BvhSystem` — compiled with feature `provider`, which needs the reference's
//! `rendiation-device-ray-tracing` crate on the path (see Cargo.toml).  This is the file a maintainer adds to make the B200 library
//! a third geometry backend next to `NaiveSahBVHSystem` (geometry/naive/mod.rs:495-610) and `HardwareInlineRayQuerySystem`
//! (wavefront_compute/mod.rs:32-36).
//!
//! The trait has two halves (shader/ray-tracing/src/api/backend.rs:120-142).  The HOST half — create / delete BLAS and TLAS,
//! `bind_tlas`, `bind_tlas_max_len` — maps one to one.  The SHADER-BUILD half (`create_comp_instance` returning an object whose
//! `build_shader` / `bind_pass` emit EDSL nodes for a wgpu compute pipeline, geometry/mod.rs:8-25) has no CUDA meaning: the
//! traversal is a hand-written kernel, not generated shader code.  Its RUNTIME contract — a `ShaderRayTraceCallStoragePayload` in,
//! an optional closest hit out, per ray — is what `B200BvhSystem::trace_closest_batch` / `trace_ray` implement for whole waves;
//! `TraceTaskImpl::device_poll` (trace_task.rs:152-360) is the call site that hands a wave over instead of polling per task.
use rendiation_device_ray_tracing::*;

use crate::{sys, B200BvhSystem, BlasGeometry};

impl Clone for B200BvhSystem {
    fn clone(&self) -> Self {
        // the reference clones an Arc; scenes here are unique owners of device memory — share it behind an Arc at the call site
        unimplemented!("wrap B200BvhSystem in an Arc: Box<dyn GPUAccelerationStructureSystemProvider> is cloned by the reference")
    }
}

impl GPUAccelerationStructureSystemProvider for B200BvhSystem {
    fn create_comp_instance(&self, _cx: &mut DeviceParallelComputeCtx) -> Box<dyn GPUAccelerationStructureSystemCompImplInstance> {
        self.commit().unwrap_or_else(|e| panic!("{e}")); // get_or_build_gpu_data (naive/mod.rs:521-536)
        unimplemented!("the EDSL shader-build half has no CUDA counterpart: waves go through trace_closest_batch / trace_ray")
    }
    fn bind_tlas_max_len(&self) -> u32 {
        B200BvhSystem::bind_tlas_max_len(self)
    }
    fn bind_tlas(&self, tlas: &[TlasHandle]) {
        let ids: Vec<crate::TlasHandle> = tlas.iter().map(|t| crate::TlasHandle(t.0)).collect();
        B200BvhSystem::bind_tlas(self, &ids)
    }
    fn create_top_level_acceleration_structure(&self, source: &[TopLevelAccelerationStructureSourceInstance]) -> TlasHandle {
        let inst: Vec<sys::rdn_instance> = source
            .iter()
            .map(|s| sys::rdn_instance {
                transform: s.transform.into(), // Mat4 is column-major a1..d4 (math/algebra/src/mat/mat4.rs:9-14)
                instance_custom_index: s.instance_custom_index,
                mask: s.mask,
                instance_shader_binding_table_record_offset: s.instance_shader_binding_table_record_offset,
                flags: s.flags as u32,
                blas_handle: s.acceleration_structure_handle.0,
            })
            .collect();
        TlasHandle(B200BvhSystem::create_top_level_acceleration_structure(self, &inst).0)
    }
    fn delete_top_level_acceleration_structure(&self, id: TlasHandle) {
        B200BvhSystem::delete_top_level_acceleration_structure(self, crate::TlasHandle(id.0))
    }
    fn create_bottom_level_acceleration_structure(&self, source: &[BottomLevelAccelerationStructureBuildSource]) -> BlasHandle {
        // Vec3<f32> is #[repr(C)] {x, y, z} (math/algebra/src/vec/vec3.rs): a slice of it is a slice of [f32; 3]
        let geoms: Vec<BlasGeometry<'_>> = source
            .iter()
            .map(|s| match &s.geometry {
                BottomLevelAccelerationStructureBuildBuffer::Triangles { positions, indices } => BlasGeometry::Triangles {
                    positions: unsafe { std::slice::from_raw_parts(positions.as_ptr() as *const [f32; 3], positions.len()) },
                    indices: indices.as_deref(),
                    flags: s.flags as u32,
                },
                BottomLevelAccelerationStructureBuildBuffer::AABBs { aabbs } => BlasGeometry::Aabbs { aabbs, flags: s.flags as u32 },
            })
            .collect();
        BlasHandle(B200BvhSystem::create_bottom_level_acceleration_structure(self, &geoms).0)
    }
    fn delete_bottom_level_acceleration_structure(&self, id: BlasHandle) {
        B200BvhSystem::delete_bottom_level_acceleration_structure(self, crate::BlasHandle(id.0))
    }
}
