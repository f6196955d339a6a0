//! Safe wrapper over `librdn_rt` (include/rdn_rt.h): the B200 drop-in for the closest-hit traversal path of rendiation.
//!
//! What it replaces in the reference (`/root/reference` = mikialex/rendiation):
//! * the host half of `GPUAccelerationStructureSystemProvider` (shader/ray-tracing/src/api/backend.rs:120-142) as implemented by
//!   `NaiveSahBVHSystem` (shader/ray-tracing/src/backend/wavefront_compute/geometry/naive/mod.rs:495-610): create / delete BLAS and
//!   TLAS, `bind_tlas`, `bind_tlas_max_len`, the lazy build;
//! * the per-ray runtime contract of `GPUAccelerationStructureSystemCompImplInvocationTraversable::traverse`
//!   (geometry/mod.rs:16-25; CPU twin `NaiveSahBvhCpu::traverse`, naive/traverse_cpu.rs:52-245) as ONE batched call,
//!   [`B200BvhSystem::trace_closest_batch`];
//! * the wave compaction of `use_compact_alive_tasks` (shader/task-graph/src/runtime/task_group.rs:220-278).
//!
//! Error behaviour: the reference panics (`panic = "abort"`, Cargo.toml:161-162) where the C side returns a status; the methods
//! that mirror trait methods keep that (`check` panics with the library's message), the additional ones return `Result`.
//!
//! The trait implementation itself lives in `provider.rs` (feature `provider`): it needs the reference's crates on the path.
//! This file has no dependency besides the `-sys` crate, so it builds wherever `nvcc` for sm_100a does.

use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::ptr::NonNull;

pub use rendiation_rt_b200_sys as sys;
pub use sys::{rdn_hit as Hit, rdn_instance as Instance, rdn_launch as Launch, rdn_ray as Ray};

#[cfg(feature = "provider")]
pub mod provider;

/// A status of the C ABI other than `RDN_OK`, with the library's thread-local message.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "rdn_rt error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

fn last_error(code: i32) -> Error {
    let message = unsafe { CStr::from_ptr(sys::rdn_rt_last_error()) }.to_string_lossy().into_owned();
    Error { code, message }
}
fn result(rc: i32) -> Result<(), Error> {
    if rc == sys::RDN_OK { Ok(()) } else { Err(last_error(rc)) }
}
/// the reference's behaviour on the Rust side of the boundary: a failed build / an invalid handle is a panic
fn check(rc: i32) {
    if let Err(e) = result(rc) {
        panic!("{e}");
    }
}

/// `BlasHandle` / `TlasHandle` of the reference (api/backend.rs:153-156): plain indices, never reused.
#[derive(Clone, Copy, Debug, PartialEq, Eq, Hash)]
pub struct BlasHandle(pub u32);
#[derive(Clone, Copy, Debug, PartialEq, Eq, Hash)]
pub struct TlasHandle(pub u32);

/// One geometry of a `BottomLevelAccelerationStructureBuildSource` (api/backend.rs:104-118), borrowed for the duration of the call
/// (the library copies its inputs, as the reference's `source.to_vec()` does).
pub enum BlasGeometry<'a> {
    Triangles { positions: &'a [[f32; 3]], indices: Option<&'a [u32]>, flags: u32 },
    /// accepted and ignored, exactly as the reference's naive builder does (naive/mod.rs:201-237)
    Aabbs { aabbs: &'a [[f32; 6]], flags: u32 },
}

/// The B200 counterpart of `NaiveSahBVHSystem`.  Interior locking is the library's (write lock for mutations, read lock for
/// traces — the discipline of naive/mod.rs:521-536), so the object is `Send + Sync` like the reference's `Arc<RwLock<..>>`.
pub struct B200BvhSystem {
    raw: NonNull<sys::rdn_rt_scene>,
}
unsafe impl Send for B200BvhSystem {}
unsafe impl Sync for B200BvhSystem {}

impl B200BvhSystem {
    /// `devices`: CUDA device ordinals; the flattened scene is replicated to each on commit and host-buffer traces are sharded over
    /// them by ray tile.  One process per GPU passes one ordinal.
    pub fn new(devices: &[i32]) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        result(unsafe { sys::rdn_rt_scene_create(devices.len() as i32, devices.as_ptr(), &mut raw) })?;
        Ok(Self { raw: NonNull::new(raw).expect("rdn_rt_scene_create returned null") })
    }
    pub fn as_raw(&self) -> *mut sys::rdn_rt_scene {
        self.raw.as_ptr()
    }

    // ---- GPUAccelerationStructureSystemProvider, host half (same names, same meaning)
    pub fn create_bottom_level_acceleration_structure(&self, source: &[BlasGeometry<'_>]) -> BlasHandle {
        let geoms: Vec<sys::rdn_blas_geometry> = source
            .iter()
            .map(|g| match g {
                BlasGeometry::Triangles { positions, indices, flags } => sys::rdn_blas_geometry {
                    positions: positions.as_ptr() as *const f32,
                    n_positions: positions.len() as u64,
                    indices: indices.map_or(std::ptr::null(), |i| i.as_ptr()),
                    n_indices: indices.map_or(0, |i| i.len() as u64),
                    flags: *flags,
                    kind: 0,
                },
                BlasGeometry::Aabbs { aabbs, flags } => sys::rdn_blas_geometry {
                    positions: aabbs.as_ptr() as *const f32,
                    n_positions: aabbs.len() as u64,
                    indices: std::ptr::null(),
                    n_indices: 0,
                    flags: *flags,
                    kind: 1,
                },
            })
            .collect();
        let mut h = 0u32;
        check(unsafe { sys::rdn_rt_blas_create(self.as_raw(), geoms.as_ptr(), geoms.len() as u32, &mut h) });
        BlasHandle(h)
    }
    pub fn delete_bottom_level_acceleration_structure(&self, id: BlasHandle) {
        check(unsafe { sys::rdn_rt_blas_destroy(self.as_raw(), id.0) })
    }
    pub fn create_top_level_acceleration_structure(&self, source: &[Instance]) -> TlasHandle {
        let mut h = 0u32;
        check(unsafe { sys::rdn_rt_tlas_create(self.as_raw(), source.as_ptr(), source.len() as u32, &mut h) });
        TlasHandle(h)
    }
    pub fn delete_top_level_acceleration_structure(&self, id: TlasHandle) {
        check(unsafe { sys::rdn_rt_tlas_destroy(self.as_raw(), id.0) })
    }
    pub fn bind_tlas(&self, tlas: &[TlasHandle]) {
        let ids: Vec<u32> = tlas.iter().map(|t| t.0).collect();
        check(unsafe { sys::rdn_rt_bind_tlas(self.as_raw(), ids.as_ptr(), ids.len() as u32) })
    }
    pub fn bind_tlas_max_len(&self) -> u32 {
        unsafe { sys::rdn_rt_bind_tlas_max_len(self.as_raw()) }
    }
    /// what `create_comp_instance` triggers through `get_or_build_gpu_data` (naive/mod.rs:521-536): build, flatten, upload, replicate
    pub fn commit(&self) -> Result<(), Error> {
        result(unsafe { sys::rdn_rt_commit(self.as_raw()) })
    }
    pub fn build_stats(&self) -> Result<sys::rdn_build_stats, Error> {
        let mut st = sys::rdn_build_stats::default();
        result(unsafe { sys::rdn_rt_scene_build_stats(self.as_raw(), &mut st) })?;
        Ok(st)
    }

    // ---- `…InvocationTraversable::traverse`, batched
    /// One closest-hit record per ray (`instance_id == u32::MAX` = the reference's `ShaderOption::is_some == false`), bit-identical to
    /// `NaiveSahBvhCpu::traverse` on the same scene and rays.  Borrowed slices are pageable memory: correct, but the copies run at a
    /// fraction of the PCIe rate — buffers that live longer than a frame belong in a [`PinnedVec`] (or [`HostRegistration`]).
    pub fn trace_closest_batch(&self, launch: Launch, rays: &[Ray]) -> Result<Vec<Hit>, Error> {
        let mut hits = Vec::<Hit>::with_capacity(rays.len());
        result(unsafe { sys::rdn_rt_trace_closest(self.as_raw(), &launch, rays.as_ptr(), rays.len() as u64, hits.as_mut_ptr()) })?;
        unsafe { hits.set_len(rays.len()) };
        Ok(hits)
    }
    /// the same into a caller-provided buffer (no allocation per call; `hits.len() >= rays.len()`)
    pub fn trace_closest_into(&self, launch: Launch, rays: &[Ray], hits: &mut [Hit]) -> Result<(), Error> {
        assert!(hits.len() >= rays.len());
        result(unsafe { sys::rdn_rt_trace_closest(self.as_raw(), &launch, rays.as_ptr(), rays.len() as u64, hits.as_mut_ptr()) })
    }
    /// reference-order walk with the reference's four visit counters (traverse_cpu.rs:37-41)
    pub fn trace_counted(&self, launch: Launch, rays: &[Ray]) -> Result<(Vec<Hit>, sys::rdn_counters), Error> {
        let mut hits = Vec::<Hit>::with_capacity(rays.len());
        let mut c = sys::rdn_counters::default();
        result(unsafe { sys::rdn_rt_trace_counted(self.as_raw(), &launch, rays.as_ptr(), rays.len() as u64, hits.as_mut_ptr(), &mut c) })?;
        unsafe { hits.set_len(rays.len()) };
        Ok((hits, c))
    }
    /// stable compaction of the active-index list (`use_stream_compaction`, parallel-compute/src/stream_compaction.rs:3-45):
    /// kept values in order, zeros behind, and the new size
    pub fn compact_u32(&self, values: &[u32], keep: &[u8]) -> Result<(Vec<u32>, u64), Error> {
        assert_eq!(values.len(), keep.len());
        let mut out = vec![0u32; values.len()];
        let mut n = 0u64;
        result(unsafe { sys::rdn_rt_compact_u32(self.as_raw(), values.as_ptr(), keep.as_ptr(), values.len() as u64, out.as_mut_ptr(), &mut n) })?;
        Ok((out, n))
    }

    // ---- device-resident path: rays and hits stay in HBM, calls are asynchronous on a CUDA stream
    /// # Safety
    /// `d_rays` / `d_hits` are device pointers on device `device_index` of this scene, 32-byte aligned, `n` records each;
    /// `overlap_previous` is the caller's promise behind `RDN_TRACE_OVERLAP_PREVIOUS` (see rdn_rt.h).
    pub unsafe fn trace_closest_device(&self, device_index: i32, launch: Launch, d_rays: *const Ray, n: u64, d_hits: *mut Hit, stream: *mut c_void,
                                       overlap_previous: bool) -> Result<(), Error> {
        let mode = sys::RDN_TRACE_AUTO | if overlap_previous { sys::RDN_TRACE_OVERLAP_PREVIOUS } else { 0 };
        result(sys::rdn_rt_trace_closest_device(self.as_raw(), device_index, &launch, d_rays, n, d_hits, stream, mode, std::ptr::null_mut()))
    }
    /// a wave whose size a previous kernel left on the device (`d_n`): no read-back between the waves of a frame
    /// # Safety
    /// as [`Self::trace_closest_device`]; `d_n` points to one `u64` on the same device.
    pub unsafe fn trace_closest_device_n(&self, device_index: i32, launch: Launch, d_rays: *const Ray, d_n: *const u64, n_max: u64, d_hits: *mut Hit,
                                         stream: *mut c_void) -> Result<(), Error> {
        result(sys::rdn_rt_trace_closest_device_n(self.as_raw(), device_index, &launch, d_rays, d_n, n_max, d_hits, stream, sys::RDN_TRACE_AUTO))
    }
    /// waits for `stream`, returns and clears the safety-net flags of the asynchronous path (`RDN_ERROR_FLAG_*`)
    pub fn poll_errors(&self, device_index: i32, stream: *mut c_void) -> Result<u32, Error> {
        let mut flags = 0u32;
        result(unsafe { sys::rdn_rt_poll_errors(self.as_raw(), device_index, stream, &mut flags) })?;
        Ok(flags)
    }
}
impl Drop for B200BvhSystem {
    fn drop(&mut self) {
        unsafe { sys::rdn_rt_scene_destroy(self.as_raw()) }
    }
}

/// A fixed-length buffer in page-locked host memory (`rdn_rt_host_alloc`): ray and hit storage that the host-buffer trace moves at
/// PCIe speed.  `T` must be plain data (the `-sys` records are).
pub struct PinnedVec<T: Copy> {
    ptr: NonNull<T>,
    len: usize,
    _own: PhantomData<T>,
}
unsafe impl<T: Copy + Send> Send for PinnedVec<T> {}
impl<T: Copy> PinnedVec<T> {
    pub fn zeroed(len: usize) -> Result<Self, Error> {
        let mut p: *mut c_void = std::ptr::null_mut();
        result(unsafe { sys::rdn_rt_host_alloc((len.max(1) * std::mem::size_of::<T>()) as u64, &mut p) })?;
        unsafe { std::ptr::write_bytes(p as *mut u8, 0, len * std::mem::size_of::<T>()) };
        Ok(Self { ptr: NonNull::new(p as *mut T).expect("rdn_rt_host_alloc returned null"), len, _own: PhantomData })
    }
}
impl<T: Copy> std::ops::Deref for PinnedVec<T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        unsafe { std::slice::from_raw_parts(self.ptr.as_ptr(), self.len) }
    }
}
impl<T: Copy> std::ops::DerefMut for PinnedVec<T> {
    fn deref_mut(&mut self) -> &mut [T] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr.as_ptr(), self.len) }
    }
}
impl<T: Copy> Drop for PinnedVec<T> {
    fn drop(&mut self) {
        unsafe { sys::rdn_rt_host_free(self.ptr.as_ptr() as *mut c_void) }
    }
}

/// Page-locks memory the caller owns for as long as the guard lives (`rdn_rt_host_register` / `_unregister`).  The borrow keeps the
/// buffer from being freed or reallocated while it is registered — the reason the library never registers caller memory by itself.
pub struct HostRegistration<'a, T> {
    ptr: *mut c_void,
    _buf: PhantomData<&'a mut [T]>,
}
impl<'a, T> HostRegistration<'a, T> {
    pub fn new(buf: &'a mut [T]) -> Result<Self, Error> {
        let ptr = buf.as_mut_ptr() as *mut c_void;
        result(unsafe { sys::rdn_rt_host_register(ptr, std::mem::size_of_val(buf) as u64) })?;
        Ok(Self { ptr, _buf: PhantomData })
    }
}
impl<T> Drop for HostRegistration<'_, T> {
    fn drop(&mut self) {
        unsafe { sys::rdn_rt_host_unregister(self.ptr) };
    }
}

#[cfg(test)]
mod tests {
    //! `cargo test` on a machine with a B200: the reference's own fixture cube (naive/test.rs:80-96) behind one instance, a pinhole
    //! grid, and the invariants any closest hit must satisfy.  (Bit-level parity with `NaiveSahBvhCpu::traverse` is held by the
    //! Python test-suite of the repository through the same C ABI.)
    use super::*;

    const CUBE_POSITION: [[f32; 3]; 8] = [
        [-0.5, -0.5, 0.5], [0.5, -0.5, 0.5], [0.5, 0.5, 0.5], [-0.5, 0.5, 0.5],
        [-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [0.5, 0.5, -0.5], [-0.5, 0.5, -0.5],
    ];
    const CUBE_INDEX: [u32; 36] = [0, 1, 2, 0, 2, 3, 1, 5, 6, 1, 6, 2, 5, 4, 7, 5, 7, 6, 4, 0, 3, 4, 3, 7, 3, 2, 6, 3, 6, 7, 4, 5, 1, 4, 1, 0];

    #[test]
    fn cube_behind_one_instance() {
        let sys_ = B200BvhSystem::new(&[0]).unwrap();
        let blas = sys_.create_bottom_level_acceleration_structure(&[BlasGeometry::Triangles {
            positions: &CUBE_POSITION, indices: Some(&CUBE_INDEX), flags: sys::RDN_GEOMETRY_FLAG_OPAQUE }]);
        let mut transform = [0f32; 16];
        for k in 0..4 { transform[5 * k] = 1.0; }
        transform[14] = -5.0; // translate(0, 0, -5): column-major, d3
        let tlas = sys_.create_top_level_acceleration_structure(&[Instance {
            transform, instance_custom_index: 7, mask: u32::MAX, instance_shader_binding_table_record_offset: 0, flags: 0, blas_handle: blas.0 }]);
        sys_.bind_tlas(&[tlas]);
        let (w, h) = (64usize, 64usize);
        let mut rays = PinnedVec::<Ray>::zeroed(w * h).unwrap();
        for j in 0..h {
            for i in 0..w {
                let (x, y) = ((i as f32 + 0.5) / w as f32 * 2.0 - 1.0, 1.0 - (j as f32 + 0.5) / h as f32 * 2.0);
                let inv = 1.0 / (x * x + y * y + 1.0f32).sqrt();
                rays[j * w + i] = Ray { ox: 0.0, oy: 0.0, oz: 0.0, tmin: 0.0, dx: x * inv, dy: y * inv, dz: -inv, tmax: 100.0 };
            }
        }
        let launch = Launch { ray_flags: sys::RDN_RAY_FLAG_CULL_BACK_FACING_TRIANGLES, cull_mask: u32::MAX, tlas_idx: 0, grid_width: w as u32 };
        let hits = sys_.trace_closest_batch(launch, &rays).unwrap();
        let centre = hits[(h / 2) * w + w / 2];
        assert_eq!(centre.instance_id, 0);
        assert_eq!(centre.instance_custom_id, 7);
        assert!((centre.t - 4.5).abs() < 0.01 && centre.primitive_id < 12);
        assert_eq!(hits[0].instance_id, sys::RDN_INVALID_ID); // the corner ray misses
        let st = sys_.build_stats().unwrap();
        assert_eq!(st.irregular_instances, 0);
    }
}
