// rdn_math.h — f32 primitives shared by the host builder/flattener and the CUDA kernels.
//
// Every expression is written in the reference's evaluation order and must never be contracted
// into FMAs: device code is compiled with -fmad=false, host code with -ffp-contract=off, and IEEE
// division / square root are used (nvcc defaults -prec-div=true -prec-sqrt=true).  That is what makes
// hit t / barycentrics bit-identical to the reference's CPU arithmetic
// (math/algebra/src/vec/vec3.rs:26-28,116-122; vec/dimension.rs:98-105; mat/mat4.rs:72-104,161-168;
//  mat/mat3.rs:37-42,103-109).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RDN_HD __host__ __device__ __forceinline__
#else
#define RDN_HD inline
#endif

namespace rdn {

struct Vec3 {
  float x, y, z;
};

RDN_HD Vec3 make_vec3(float x, float y, float z) { return Vec3{x, y, z}; }
RDN_HD Vec3 operator+(Vec3 a, Vec3 b) { return Vec3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RDN_HD Vec3 operator-(Vec3 a, Vec3 b) { return Vec3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RDN_HD Vec3 operator*(Vec3 a, Vec3 b) { return Vec3{a.x * b.x, a.y * b.y, a.z * b.z}; }
RDN_HD Vec3 operator*(Vec3 a, float s) { return Vec3{a.x * s, a.y * s, a.z * s}; }
RDN_HD Vec3 operator/(Vec3 a, float s) { return Vec3{a.x / s, a.y / s, a.z / s}; }
RDN_HD float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RDN_HD Vec3 cross(Vec3 a, Vec3 b) {
  return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// fminf / fmaxf (a NaN operand loses).  Host compilers keep them as calls into libm unless given -ffinite-math-only — a third of the
// host builder's time went there — so the host spells the same selection out (glibc's own: x if x <= y, y if y < x, else the non-NaN).
RDN_HD float min_f32(float a, float b) {
#ifdef __CUDA_ARCH__
  return fminf(a, b);
#else
  return a <= b ? a : (b < a ? b : (b != b ? a : b));
#endif
}
RDN_HD float max_f32(float a, float b) {
#ifdef __CUDA_ARCH__
  return fmaxf(a, b);
#else
  return a >= b ? a : (b > a ? b : (b != b ? a : b));
#endif
}
RDN_HD Vec3 vmin(Vec3 a, Vec3 b) { return Vec3{min_f32(a.x, b.x), min_f32(a.y, b.y), min_f32(a.z, b.z)}; }
RDN_HD Vec3 vmax(Vec3 a, Vec3 b) { return Vec3{max_f32(a.x, b.x), max_f32(a.y, b.y), max_f32(a.z, b.z)}; }
RDN_HD float length(Vec3 a) { return sqrtf(dot(a, a)); }
// InnerProductSpace::normalize: unchanged when the squared length is not > 0
RDN_HD Vec3 normalize(Vec3 a) {
  float mag_sq = dot(a, a);
  if (mag_sq > 0.0f) {
    float inv_sqrt = 1.0f / sqrtf(mag_sq);
    return a * inv_sqrt;
  }
  return a;
}

struct Box3 {
  Vec3 min, max;
};
RDN_HD Box3 box_empty() { return Box3{Vec3{INFINITY, INFINITY, INFINITY}, Vec3{-INFINITY, -INFINITY, -INFINITY}}; }
RDN_HD void expand(Box3 &b, Vec3 p) { b.min = vmin(b.min, p); b.max = vmax(b.max, p); }
RDN_HD void expand(Box3 &b, const Box3 &o) { b.min = vmin(b.min, o.min); b.max = vmax(b.max, o.max); }

// column-major, fields named as the reference's Mat4 (a* = column 0 ... d* = column 3)
struct Mat4 {
  float a1, a2, a3, a4, b1, b2, b3, b4, c1, c2, c3, c4, d1, d2, d3, d4;
};

}  // namespace rdn
