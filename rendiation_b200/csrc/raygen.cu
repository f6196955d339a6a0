// raygen.cu — the step on either side of the traversal: primary-ray generation and the
// closest-hit -> bounce-ray step, on the device, so a frame never round-trips rays or hits through the host.
//
//   k_gen_pinhole_rays   the reference's test pinhole grid (shader/ray-tracing/.../naive/test.rs:259-264) with the
//                        sub-pixel jitter / aspect correction of the BASELINE configs
//   k_gen_camera_rays    DefaultRtxCameraInvocation::generate_ray (scene/rendering/gpu-ray-tracing/src/camera.rs:66-98):
//                        uv = pixel/size + sampler.next_2d()/size, unprojected through view_projection_inv
//                        (shader_uv_space_to_render_space, shader/library/src/lib.rs:18-28); the sampler is
//                        PCGRandomSampler seeded by xxhash32(pixel.x, pixel.y, sample_index) (sampler.rs:11-72)
//   k_mark_hits          keep flag per ray (hit <=> instance_id != INVALID) feeding the wavefront compaction (compact.cu)
//   k_ao_scatter / k_ao_accumulate   the AO frame's payload and running mean (feature/ao.rs:187-232)
//   k_gen_bounce_rays    one ray per surviving hit: origin = hit_world_position (api/ctx.rs:209-211: origin + dir * t),
//                        geometric normal = normalize(normal_mat * (pa-pb) x (pa-pc)) turned towards the ray origin
//                        (bindless_mesh_bridge.rs:103-114), direction either
//                          mode 0: cosine_sample_hemisphere_in_dir(normal, (van_der_corput, sobol)(source ray index))
//                                  (math/statistics/src/distribution_map.rs:10-58, sampling/sobol.rs:40-68; fixed scrambles,
//                                  SURVEY.md §8d C3), or
//                          mode 1: tbn(normal) * sample_hemisphere_cos(hammersley_2d(sample_index, 256))
//                                  (feature/ao.rs:249-284, shader/library/src/sampling.rs:33-83)
//
// Arithmetic is plain f32 in the written order, compiled with -fmad=false like the traversal; sinf/cosf are CUDA's,
// so directions agree with a numpy restatement to a few ulp, not bit for bit (the traversal parity is then checked on
// exactly the rays these kernels produced).
#include <cuda_runtime.h>

#include "kernels.h"
#include "rdn_math.h"

namespace rdn {

namespace {

__device__ __forceinline__ void store_ray(rdn_ray *dst, Vec3 o, float tmin, Vec3 d, float tmax) {
  float4 *p = reinterpret_cast<float4 *>(dst);
  p[0] = make_float4(o.x, o.y, o.z, tmin);
  p[1] = make_float4(d.x, d.y, d.z, tmax);
}

__global__ void __launch_bounds__(256) k_gen_pinhole_rays(const rdn_pinhole P, rdn_ray *__restrict__ rays) {
  const uint64_t n = static_cast<uint64_t>(P.rect_w) * P.rect_h;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t ry = static_cast<uint32_t>(k / P.rect_w), rx = static_cast<uint32_t>(k - static_cast<uint64_t>(ry) * P.rect_w);
    const float i = static_cast<float>(P.rect_x + rx), j = static_cast<float>(P.rect_y + ry);
    float x = (i + P.jitter_x) / static_cast<float>(P.width) * 2.0f - 1.0f;
    const float y = 1.0f - (j + P.jitter_y) / static_cast<float>(P.height) * 2.0f;
    if (P.aspect != 1.0f) x = x * P.aspect;
    const Vec3 o = {P.origin[0], P.origin[1], P.origin[2]};
    const Vec3 target = {x, y, -1.0f};
    store_ray(rays + k, o, P.tmin, normalize(target - o), P.tmax);
  }
}

// many rectangles in one launch: blockIdx.y = descriptor, rays of descriptor k start at offsets[k]
__global__ void __launch_bounds__(256) k_gen_pinhole_rays_batch(const rdn_pinhole *__restrict__ params, const uint64_t *__restrict__ offsets,
                                                                rdn_ray *__restrict__ rays) {
  const rdn_pinhole P = params[blockIdx.y];
  rdn_ray *dst = rays + offsets[blockIdx.y];
  const uint64_t n = static_cast<uint64_t>(P.rect_w) * P.rect_h;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t ry = static_cast<uint32_t>(k / P.rect_w), rx = static_cast<uint32_t>(k - static_cast<uint64_t>(ry) * P.rect_w);
    const float i = static_cast<float>(P.rect_x + rx), j = static_cast<float>(P.rect_y + ry);
    float x = (i + P.jitter_x) / static_cast<float>(P.width) * 2.0f - 1.0f;
    const float y = 1.0f - (j + P.jitter_y) / static_cast<float>(P.height) * 2.0f;
    if (P.aspect != 1.0f) x = x * P.aspect;
    const Vec3 o = {P.origin[0], P.origin[1], P.origin[2]};
    const Vec3 target = {x, y, -1.0f};
    store_ray(dst + k, o, P.tmin, normalize(target - o), P.tmax);
  }
}

// ---- sampler.rs:11-72
__device__ __forceinline__ uint32_t xxhash32(uint32_t px, uint32_t py, uint32_t pz) {
  const uint32_t p0 = 2246822519u, p1 = 3266489917u, p2 = 668265263u, p3 = 374761393u;
  uint32_t h = pz + p3 + px * p1;
  h = p2 * ((h << 17) | (h >> 15));
  h = h + py * p1;
  h = p2 * ((h << 17) | (h >> 15));
  h = p0 * (h ^ (h >> 15));
  h = p1 * (h ^ (h >> 13));
  return h ^ (h >> 16);
}
__device__ __forceinline__ float pcg_next(uint32_t &state) {
  const uint32_t prev = state * 747796405u + 2891336453u;
  const uint32_t word = ((prev >> ((prev >> 28) + 4u)) ^ prev) * 277803737u;
  state = prev;
  const uint32_t r = (word >> 22) ^ word;
  return __uint_as_float(0x3f800000u | (r >> 9)) - 1.0f;
}

__global__ void __launch_bounds__(256) k_gen_camera_rays(const rdn_camera P, rdn_ray *__restrict__ rays) {
  const uint64_t n = static_cast<uint64_t>(P.rect_w) * P.rect_h;
  const float *m = P.view_projection_inv;  // column-major a1..d4
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t ry = static_cast<uint32_t>(k / P.rect_w), rx = static_cast<uint32_t>(k - static_cast<uint64_t>(ry) * P.rect_w);
    const uint32_t px = P.rect_x + rx, py = P.rect_y + ry;
    uint32_t state = xxhash32(px, py, P.sample_index);
    const float s0 = pcg_next(state), s1 = pcg_next(state);
    const float fw = static_cast<float>(P.width), fh = static_cast<float>(P.height);
    const float u = static_cast<float>(px) / fw + s0 / fw;
    const float v = static_cast<float>(py) / fh + s1 / fh;
    const float nx = (u * 2.0f - 1.0f) * 1.0f, ny = (v * 2.0f - 1.0f) * -1.0f, nz = P.ndc_depth, nw = 1.0f;
    // Mat4 * Vec4 (mat4.rs:161-168)
    const float x = nx * m[0] + ny * m[4] + nz * m[8] + nw * m[12];
    const float y = nx * m[1] + ny * m[5] + nz * m[9] + nw * m[13];
    const float z = nx * m[2] + ny * m[6] + nz * m[10] + nw * m[14];
    const float w = nx * m[3] + ny * m[7] + nz * m[11] + nw * m[15];
    const Vec3 target = {x / w, y / w, z / w};
    const Vec3 o = {P.world_position[0], P.world_position[1], P.world_position[2]};
    store_ray(rays + k, o, P.tmin, normalize(target - o), P.tmax);
  }
}

__global__ void __launch_bounds__(256) k_mark_hits(const rdn_hit *__restrict__ hits, uint64_t n, uint8_t *__restrict__ keep,
                                                   uint32_t *__restrict__ iota) {
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    keep[k] = hits[k].instance_id != RDN_INVALID_ID ? 1 : 0;
    iota[k] = static_cast<uint32_t>(k);
  }
}

// ---- the AO frame's accumulation (feature/ao.rs:187-232): payload = 1 where the primary ray misses (miss shader) or its AO test
// ray misses, 0 where the AO test ray hits anything (secondary closest-hit shader); the running mean over the samples so far
// is updated in place, frozen once max_sample samples are in.  One thread per pixel; the AO ray of pixel p (if any) is the
// k-th compacted bounce ray with src_index[k] == p, found through a scatter pass.
__global__ void __launch_bounds__(256) k_ao_scatter(const rdn_hit *__restrict__ secondary_hits, const uint32_t *__restrict__ src_index,
                                                    const uint64_t *__restrict__ n_secondary, float *__restrict__ payload) {
  const uint64_t n = *n_secondary;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    payload[src_index[k]] = secondary_hits[k].instance_id != RDN_INVALID_ID ? 0.0f : 1.0f;
}
__global__ void __launch_bounds__(256) k_ao_accumulate(float *__restrict__ payload_then_unused, uint64_t n, uint32_t sample_count, uint32_t max_sample,
                                                       float *__restrict__ ao_buffer) {
  const float previous_sample_count = static_cast<float>(sample_count);
  const float all_sample_count = previous_sample_count + 1.0f;
  for (uint64_t p = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; p < n; p += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const float payload = payload_then_unused[p];
    payload_then_unused[p] = 1.0f;  // re-armed: "miss" is the default of the next sample
    if (sample_count < max_sample) ao_buffer[p] = (ao_buffer[p] * previous_sample_count + payload) / all_sample_count;
  }
}
__global__ void __launch_bounds__(256) k_fill_f32(float *__restrict__ dst, uint64_t n, float v) {
  for (uint64_t p = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; p < n; p += static_cast<uint64_t>(gridDim.x) * blockDim.x) dst[p] = v;
}

// ---- low-discrepancy points with fixed scrambles (sampling/sobol.rs:40-68)
__device__ __forceinline__ float to_unit_float24(uint32_t bits) {
  const float v = static_cast<float>((bits >> 8) & 0xffffffu) / 16777216.0f;
  return fminf(v, 1.0f - 1.1920929e-07f);
}
__device__ __forceinline__ float van_der_corput(uint32_t n, uint32_t scramble) {
  n = (n >> 16) | (n << 16);
  n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
  n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
  n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
  n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
  return to_unit_float24(n ^ scramble);
}
__device__ __forceinline__ float sobol2(uint32_t n, uint32_t scramble) {
  uint32_t s = scramble;
  for (uint32_t i = 1u << 31; n != 0; n >>= 1, i ^= i >> 1)
    if (n & 1u) s ^= i;
  return to_unit_float24(s);
}
__device__ __forceinline__ float radical_inverse_vdc(uint32_t bits) {  // shader/library/src/sampling.rs:43-52
  bits = (bits << 16) | (bits >> 16);
  bits = ((bits & 0x55555555u) << 1) | ((bits & 0xAAAAAAAAu) >> 1);
  bits = ((bits & 0x33333333u) << 2) | ((bits & 0xCCCCCCCCu) >> 2);
  bits = ((bits & 0x0F0F0F0Fu) << 4) | ((bits & 0xF0F0F0F0u) >> 4);
  bits = ((bits & 0x00FF00FFu) << 8) | ((bits & 0xFF00FF00u) >> 8);
  return static_cast<float>(bits) * 2.3283064e-10f;
}

// distribution_map.rs:10-58: concentric disk -> hemisphere around `dir`
__device__ __forceinline__ Vec3 cosine_sample_hemisphere_in_dir(Vec3 dir, float s0, float s1) {
  const float ux = s0 * 2.0f - 1.0f, uy = s1 * 2.0f - 1.0f;
  float dx = 0.f, dy = 0.f;
  if (!(ux == 0.0f && uy == 0.0f)) {
    float r, theta;
    if (fabsf(ux) > fabsf(uy)) { r = ux; theta = 0.78539816339744830962f * (uy / ux); }
    else { r = uy; theta = 1.57079632679489661923f - 0.78539816339744830962f * (ux / uy); }
    dx = cosf(theta) * r;
    dy = sinf(theta) * r;
  }
  const float z = sqrtf(fmaxf(0.0f, 1.0f - dx * dx - dy * dy));
  const Vec3 up_y = {0.f, 1.f, 0.f};
  const Vec3 left = normalize(cross(up_y, dir));
  const Vec3 up = cross(left, dir);
  const float xy_r = sqrtf(dx * dx + dy * dy);
  if (xy_r == 0.0f) return dir;
  const float cos_phi = dx / xy_r, sin_phi = dy / xy_r;
  const Vec3 out = left * (xy_r * cos_phi) + up * (xy_r * sin_phi) + dir * z;
  return normalize(out);
}

// shader/library/src/sampling.rs:66-83 (tbn, Pixar orthonormal basis) and :55-62 (sample_hemisphere_cos)
__device__ __forceinline__ Vec3 ao_direction(Vec3 nrm, uint32_t sample_index, uint32_t max_sample) {
  const float hx = static_cast<float>(sample_index) / static_cast<float>(max_sample), hy = radical_inverse_vdc(sample_index);
  const float phi = 6.28318530717958647692f * hy;
  const float cos_theta = sqrtf(1.0f - hx);
  const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
  const Vec3 local = {cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta};
  const float sign = nrm.z < 0.0f ? -1.0f : 1.0f;
  const float a = -1.0f / (sign + nrm.z);
  const float b = nrm.x * nrm.y * a;
  const Vec3 tangent = normalize(Vec3{1.0f + sign * nrm.x * nrm.x * a, sign * b, -sign * nrm.x});
  const Vec3 bi_tangent = normalize(Vec3{b, sign + nrm.y * nrm.y * a, -nrm.y});
  // Mat3(columns tangent, bi_tangent, normal) * local (mat3.rs:103-109)
  return Vec3{tangent.x * local.x + bi_tangent.x * local.y + nrm.x * local.z,
              tangent.y * local.x + bi_tangent.y * local.y + nrm.y * local.z,
              tangent.z * local.x + bi_tangent.z * local.y + nrm.z * local.z};
}

// offset_ray_hit (scene/rendering/gpu-ray-tracing/src/ray_util.rs:6-40): per component, move the position by
// int(256 * n) units in the last place away from the surface (towards the normal), or by n / 65536 near the origin
__device__ __forceinline__ float offset_component(float p, float n) {
  const int of_i = static_cast<int>(n * 256.0f);  // into_i32: truncation
  const float p_i = __int_as_float(__float_as_int(p) + (p < 0.0f ? -of_i : of_i));
  return fabsf(p) < (1.0f / 32.0f) ? p + (1.0f / 65536.0f) * n : p_i;
}
__device__ __forceinline__ Vec3 offset_ray_hit(Vec3 p, Vec3 n) {
  return Vec3{offset_component(p.x, n.x), offset_component(p.y, n.y), offset_component(p.z, n.z)};
}

// SCATTER (a stage of rdn_rt_trace_ray): ray k is written to slot src of rays_out and announced in spawn[src]; its low-discrepancy
// index is the launch index of the source ray.  Otherwise (the compacting bounce step): ray k goes to rays_out[k].
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_gen_bounce_rays(const SceneDev S, const rdn_bounce P, const rdn_ray *__restrict__ rays_in,
                                                         const rdn_hit *__restrict__ hits, const uint32_t *__restrict__ src_index,
                                                         const uint64_t *__restrict__ n_src, rdn_ray *__restrict__ rays_out,
                                                         const uint32_t *__restrict__ launch_index, uint8_t *__restrict__ spawn) {
  const uint64_t n = *n_src;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t src = src_index[k];
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays_in + src));
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays_in + src) + 1);
    const float4 h0 = __ldg(reinterpret_cast<const float4 *>(hits + src));      // t u v primitive_id
    const float4 h1 = __ldg(reinterpret_cast<const float4 *>(hits + src) + 1);  // geometry_id instance_id custom hit_kind
    const Vec3 ro = {r0.x, r0.y, r0.z}, rd = {r1.x, r1.y, r1.z};
    const Vec3 pos = ro + rd * h0.x;  // hit_world_position
    // geometric normal: the slot of (instance, geometry, primitive) holds e1 = v1-v0, e2 = v2-v0 in object space, and
    // (pa-pb) x (pa-pc) == e1 x e2 exactly (negation is exact in IEEE arithmetic)
    const uint32_t inst = __float_as_uint(h1.y), geom = __float_as_uint(h1.x), prim = __float_as_uint(h0.w);
    const InstanceRecord *rec = S.instances + inst;
    const uint32_t blas = rec->blas;
    // AABB geometries own no GeometryMeta, so the record of geometry `geom` is searched in the BLAS's (short) range
    uint32_t gi = S.blas_meta[blas].tri_root_range[0];
    const uint32_t gi_end = S.blas_meta[blas].tri_root_range[1];
    while (gi + 1u < gi_end && S.geometry_meta[gi].geometry_idx != geom) ++gi;
    const uint32_t slot = S.prim_to_slot[S.geometry_meta[gi].primitive_start + prim];
    const float4 qe1 = __ldg(reinterpret_cast<const float4 *>(S.triangles + slot) + 2);
    const float4 qe2 = __ldg(reinterpret_cast<const float4 *>(S.triangles + slot) + 3);
    const Vec3 c = cross(Vec3{qe1.x, qe1.y, qe1.z}, Vec3{qe2.x, qe2.y, qe2.z});
    // normal_mat = transpose(mat3(world_to_object)): row k of the product is column k of world_to_object dotted with c
    const float *wi = rec->transform_inv;
    Vec3 g = normalize(Vec3{wi[0] * c.x + wi[1] * c.y + wi[2] * c.z, wi[4] * c.x + wi[5] * c.y + wi[6] * c.z,
                            wi[8] * c.x + wi[9] * c.y + wi[10] * c.z});
    if (dot(ro - pos, g) < 0.0f) g = Vec3{-g.x, -g.y, -g.z};
    Vec3 dir;
    float tmax = P.tmax;
    if (P.mode == 0) {
      const uint32_t sample = (SCATTER && launch_index ? launch_index[src] : src) + P.index_base;
      dir = cosine_sample_hemisphere_in_dir(g, van_der_corput(sample, P.scramble0), sobol2(sample, P.scramble1));
    } else if (P.mode == 1) {
      dir = ao_direction(g, P.sample_index, P.max_sample);
    } else {
      // PointLight::importance_sampling_light_impl (lighting_bridge.rs:86-94), from the un-offset hit position (ray_hit.rs:20-23)
      const Vec3 to_light = Vec3{P.target[0], P.target[1], P.target[2]} - pos;
      const float distance = length(to_light);
      dir = to_light / distance;
      tmax = distance;
    }
    const Vec3 origin = (P.flags & RDN_BOUNCE_OFFSET_ORIGIN) ? offset_ray_hit(pos, g) : pos;
    store_ray(rays_out + (SCATTER ? src : k), origin, P.tmin, dir, tmax);
    if (SCATTER) spawn[src] = 1;
  }
}

int grid_for(uint64_t n, int block) {
  const uint64_t b = (n + block - 1) / block;
  return static_cast<int>(b < 1 ? 1 : (b > 148ull * 32 ? 148ull * 32 : b));
}

}  // namespace

void launch_gen_pinhole_rays(const rdn_pinhole &p, rdn_ray *d_rays, cudaStream_t stream) {
  const uint64_t n = static_cast<uint64_t>(p.rect_w) * p.rect_h;
  if (n) k_gen_pinhole_rays<<<grid_for(n, 256), 256, 0, stream>>>(p, d_rays);
}
void launch_gen_pinhole_rays_batch(const rdn_pinhole *d_params, const uint64_t *d_offsets, uint32_t n_params, uint64_t max_rays_per_param,
                                   rdn_ray *d_rays, cudaStream_t stream) {
  if (n_params == 0 || max_rays_per_param == 0) return;
  const uint64_t bx = (max_rays_per_param + 255) / 256;
  const dim3 grid(static_cast<unsigned>(bx < 64 ? bx : 64), n_params);  // a few CTAs per rectangle, grid-stride inside
  k_gen_pinhole_rays_batch<<<grid, 256, 0, stream>>>(d_params, d_offsets, d_rays);
}
void launch_gen_camera_rays(const rdn_camera &p, rdn_ray *d_rays, cudaStream_t stream) {
  const uint64_t n = static_cast<uint64_t>(p.rect_w) * p.rect_h;
  if (n) k_gen_camera_rays<<<grid_for(n, 256), 256, 0, stream>>>(p, d_rays);
}
void launch_mark_hits(const rdn_hit *d_hits, uint64_t n, uint8_t *d_keep, uint32_t *d_iota, cudaStream_t stream) {
  if (n) k_mark_hits<<<grid_for(n, 256), 256, 0, stream>>>(d_hits, n, d_keep, d_iota);
}
void launch_fill_f32(float *d_dst, uint64_t n, float v, cudaStream_t stream) {
  if (n) k_fill_f32<<<grid_for(n, 256), 256, 0, stream>>>(d_dst, n, v);
}
void launch_ao_accumulate(const rdn_hit *d_secondary_hits, const uint32_t *d_src_index, const uint64_t *d_n_secondary, uint64_t n_pixels,
                          uint32_t sample_count, uint32_t max_sample, float *d_payload, float *d_ao_buffer, cudaStream_t stream) {
  if (n_pixels == 0) return;
  k_ao_scatter<<<grid_for(n_pixels, 256), 256, 0, stream>>>(d_secondary_hits, d_src_index, d_n_secondary, d_payload);
  k_ao_accumulate<<<grid_for(n_pixels, 256), 256, 0, stream>>>(d_payload, n_pixels, sample_count, max_sample, d_ao_buffer);
}
void launch_gen_bounce_rays(const SceneDev &scene, const rdn_bounce &p, const rdn_ray *d_rays_in, const rdn_hit *d_hits,
                            const uint32_t *d_src_index, const uint64_t *d_n_src, uint64_t n_max, rdn_ray *d_rays_out, cudaStream_t stream) {
  if (n_max) k_gen_bounce_rays<false><<<grid_for(n_max, 256), 256, 0, stream>>>(scene, p, d_rays_in, d_hits, d_src_index, d_n_src, d_rays_out, nullptr, nullptr);
}
void launch_stage_bounce_rays(const SceneDev &scene, const rdn_bounce &p, const rdn_ray *d_rays_in, const rdn_hit *d_hits, const uint32_t *d_tasks,
                              const uint64_t *d_count, uint64_t n_max, const uint32_t *d_launch_index, rdn_ray *d_rays_out, uint8_t *d_spawn,
                              cudaStream_t stream) {
  if (n_max) k_gen_bounce_rays<true><<<grid_for(n_max, 256), 256, 0, stream>>>(scene, p, d_rays_in, d_hits, d_tasks, d_count, d_rays_out, d_launch_index, d_spawn);
}
void launch_ao_resolve(float *d_payload, uint64_t n_pixels, uint32_t sample_count, uint32_t max_sample, float *d_ao_buffer, cudaStream_t stream) {
  if (n_pixels) k_ao_accumulate<<<grid_for(n_pixels, 256), 256, 0, stream>>>(d_payload, n_pixels, sample_count, max_sample, d_ao_buffer);
}

}  // namespace rdn
