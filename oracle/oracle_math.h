/* ORACLE — TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * f32 vector / matrix / box primitives restated op-for-op from the reference's
 * math crates so every rounding happens in the reference's order.  Build with
 * -ffp-contract=off (no FMA contraction) — Rust never contracts a*b+c.
 *
 * Follows (all under /root/reference):
 *   math/algebra/src/vec/vec3.rs:26-28   dot_impl  x*bx + y*by + z*bz (left to right)
 *   math/algebra/src/vec/vec3.rs:90-96   max_channel / min_channel
 *   math/algebra/src/vec/vec3.rs:116-122 cross
 *   math/algebra/src/vec/dimension.rs:53-60   component-wise min / max (f32::min/max = IEEE minNum/maxNum)
 *   math/algebra/src/vec/dimension.rs:98-105  normalize: v * (1/sqrt(len2)) iff len2 > 0, else v unchanged
 *   math/algebra/src/mat/mat4.rs:9-14    column-major fields a1..d4
 *   math/algebra/src/mat/mat4.rs:40-104  det (24 terms) and cofactor inverse with inv_det = 1/det
 *   math/algebra/src/mat/mat4.rs:140-168 Mat4*Vec3 (affine + /w) and Mat4*Vec4
 *   math/algebra/src/mat/mat4.rs:176-204 Mat4*Mat4
 *   math/algebra/src/mat/mat4.rs:303-384 rotate_x/y/z, scale, translate
 *   math/algebra/src/mat/mat3.rs:37-42,103-109  Mat3 det, Mat3*Vec3
 *   math/geometry/src/hyper_aabb.rs:20-50     Box3 empty / expand
 *   math/geometry/src/dimension3/box3.rs:5-38,91-93,117-133  surface area, apply_matrix, center, longest_axis
 */
#ifndef RDN_ORACLE_MATH_H
#define RDN_ORACLE_MATH_H
#include <math.h>
#include <stdint.h>

typedef struct { float x, y, z; } ov3;
typedef struct { float x, y, z, w; } ov4;
/* column-major: a* is column 0, d* is column 3 (translation lives in d1,d2,d3) */
typedef struct { float a1,a2,a3,a4, b1,b2,b3,b4, c1,c2,c3,c4, d1,d2,d3,d4; } om4;
typedef struct { float a1,a2,a3, b1,b2,b3, c1,c2,c3; } om3;
typedef struct { ov3 min, max; } obox;

static inline ov3 ov3_new(float x, float y, float z) { ov3 r = {x, y, z}; return r; }
static inline ov3 ov3_add(ov3 a, ov3 b) { return ov3_new(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline ov3 ov3_sub(ov3 a, ov3 b) { return ov3_new(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline ov3 ov3_mul(ov3 a, ov3 b) { return ov3_new(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline ov3 ov3_div(ov3 a, ov3 b) { return ov3_new(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline ov3 ov3_scale(ov3 a, float s) { return ov3_new(a.x * s, a.y * s, a.z * s); }
static inline ov3 ov3_divs(ov3 a, float s) { return ov3_new(a.x / s, a.y / s, a.z / s); }
static inline float ov3_dot(ov3 a, ov3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline ov3 ov3_cross(ov3 a, ov3 b) {
  return ov3_new(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline ov3 ov3_min(ov3 a, ov3 b) { return ov3_new(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline ov3 ov3_max(ov3 a, ov3 b) { return ov3_new(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline float ov3_max_channel(ov3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
static inline float ov3_min_channel(ov3 a) { return fminf(fminf(a.x, a.y), a.z); }
static inline float ov3_length(ov3 a) { return sqrtf(ov3_dot(a, a)); }
static inline ov3 ov3_normalize(ov3 a) {
  float mag_sq = ov3_dot(a, a);
  if (mag_sq > 0.0f) {
    float inv_sqrt = 1.0f / sqrtf(mag_sq);
    return ov3_scale(a, inv_sqrt);
  }
  return a;
}
/* Rust f32::signum: +0 -> 1, -0 -> -1, NaN -> NaN */
static inline float of_signum(float v) { return isnan(v) ? v : copysignf(1.0f, v); }

static inline om4 om4_identity(void) {
  om4 m = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
  return m;
}
static inline ov4 om4_mul_v4(om4 m, ov4 v) {
  ov4 r;
  r.x = v.x * m.a1 + v.y * m.b1 + v.z * m.c1 + v.w * m.d1;
  r.y = v.x * m.a2 + v.y * m.b2 + v.z * m.c2 + v.w * m.d2;
  r.z = v.x * m.a3 + v.y * m.b3 + v.z * m.c3 + v.w * m.d3;
  r.w = v.x * m.a4 + v.y * m.b4 + v.z * m.c4 + v.w * m.d4;
  return r;
}
/* Mat4 * Vec3 : expand_with_one, multiply, divide xyz by w */
static inline ov3 om4_mul_v3(om4 m, ov3 v) {
  ov4 e = {v.x, v.y, v.z, 1.0f};
  ov4 r = om4_mul_v4(m, e);
  return ov3_new(r.x / r.w, r.y / r.w, r.z / r.w);
}
static inline om4 om4_mul(om4 a, om4 m) {
  om4 r;
  r.a1 = a.a1 * m.a1 + a.b1 * m.a2 + a.c1 * m.a3 + a.d1 * m.a4;
  r.a2 = a.a2 * m.a1 + a.b2 * m.a2 + a.c2 * m.a3 + a.d2 * m.a4;
  r.a3 = a.a3 * m.a1 + a.b3 * m.a2 + a.c3 * m.a3 + a.d3 * m.a4;
  r.a4 = a.a4 * m.a1 + a.b4 * m.a2 + a.c4 * m.a3 + a.d4 * m.a4;
  r.b1 = a.a1 * m.b1 + a.b1 * m.b2 + a.c1 * m.b3 + a.d1 * m.b4;
  r.b2 = a.a2 * m.b1 + a.b2 * m.b2 + a.c2 * m.b3 + a.d2 * m.b4;
  r.b3 = a.a3 * m.b1 + a.b3 * m.b2 + a.c3 * m.b3 + a.d3 * m.b4;
  r.b4 = a.a4 * m.b1 + a.b4 * m.b2 + a.c4 * m.b3 + a.d4 * m.b4;
  r.c1 = a.a1 * m.c1 + a.b1 * m.c2 + a.c1 * m.c3 + a.d1 * m.c4;
  r.c2 = a.a2 * m.c1 + a.b2 * m.c2 + a.c2 * m.c3 + a.d2 * m.c4;
  r.c3 = a.a3 * m.c1 + a.b3 * m.c2 + a.c3 * m.c3 + a.d3 * m.c4;
  r.c4 = a.a4 * m.c1 + a.b4 * m.c2 + a.c4 * m.c3 + a.d4 * m.c4;
  r.d1 = a.a1 * m.d1 + a.b1 * m.d2 + a.c1 * m.d3 + a.d1 * m.d4;
  r.d2 = a.a2 * m.d1 + a.b2 * m.d2 + a.c2 * m.d3 + a.d2 * m.d4;
  r.d3 = a.a3 * m.d1 + a.b3 * m.d2 + a.c3 * m.d3 + a.d3 * m.d4;
  r.d4 = a.a4 * m.d1 + a.b4 * m.d2 + a.c4 * m.d3 + a.d4 * m.d4;
  return r;
}
static inline float om4_det(om4 m) {
  return m.a1 * m.b2 * m.c3 * m.d4
    - m.a1 * m.b2 * m.c4 * m.d3
    + m.a1 * m.b3 * m.c4 * m.d2
    - m.a1 * m.b3 * m.c2 * m.d4
    + m.a1 * m.b4 * m.c2 * m.d3
    - m.a1 * m.b4 * m.c3 * m.d2
    - m.a2 * m.b3 * m.c4 * m.d1
    + m.a2 * m.b3 * m.c1 * m.d4
    - m.a2 * m.b4 * m.c1 * m.d3
    + m.a2 * m.b4 * m.c3 * m.d1
    - m.a2 * m.b1 * m.c3 * m.d4
    + m.a2 * m.b1 * m.c4 * m.d3
    + m.a3 * m.b4 * m.c1 * m.d2
    - m.a3 * m.b4 * m.c2 * m.d1
    + m.a3 * m.b1 * m.c2 * m.d4
    - m.a3 * m.b1 * m.c4 * m.d2
    + m.a3 * m.b2 * m.c4 * m.d1
    - m.a3 * m.b2 * m.c1 * m.d4
    - m.a4 * m.b1 * m.c2 * m.d3
    + m.a4 * m.b1 * m.c3 * m.d2
    - m.a4 * m.b2 * m.c3 * m.d1
    + m.a4 * m.b2 * m.c1 * m.d3
    - m.a4 * m.b3 * m.c1 * m.d2
    + m.a4 * m.b3 * m.c2 * m.d1;
}
/* inverse_or_identity (mat/dimension.rs:18-20 over mat4.rs:72-104) */
static inline om4 om4_inverse_or_identity(om4 m) {
  float det = om4_det(m);
  if (det == 0.0f) return om4_identity();
  float inv_det = 1.0f / det;
  float n = -inv_det;
  om4 r;
  r.a1 = inv_det * (m.b2 * (m.c3 * m.d4 - m.c4 * m.d3) + m.b3 * (m.c4 * m.d2 - m.c2 * m.d4) + m.b4 * (m.c2 * m.d3 - m.c3 * m.d2));
  r.a2 = n * (m.a2 * (m.c3 * m.d4 - m.c4 * m.d3) + m.a3 * (m.c4 * m.d2 - m.c2 * m.d4) + m.a4 * (m.c2 * m.d3 - m.c3 * m.d2));
  r.a3 = inv_det * (m.a2 * (m.b3 * m.d4 - m.b4 * m.d3) + m.a3 * (m.b4 * m.d2 - m.b2 * m.d4) + m.a4 * (m.b2 * m.d3 - m.b3 * m.d2));
  r.a4 = n * (m.a2 * (m.b3 * m.c4 - m.b4 * m.c3) + m.a3 * (m.b4 * m.c2 - m.b2 * m.c4) + m.a4 * (m.b2 * m.c3 - m.b3 * m.c2));
  r.b1 = n * (m.b1 * (m.c3 * m.d4 - m.c4 * m.d3) + m.b3 * (m.c4 * m.d1 - m.c1 * m.d4) + m.b4 * (m.c1 * m.d3 - m.c3 * m.d1));
  r.b2 = inv_det * (m.a1 * (m.c3 * m.d4 - m.c4 * m.d3) + m.a3 * (m.c4 * m.d1 - m.c1 * m.d4) + m.a4 * (m.c1 * m.d3 - m.c3 * m.d1));
  r.b3 = n * (m.a1 * (m.b3 * m.d4 - m.b4 * m.d3) + m.a3 * (m.b4 * m.d1 - m.b1 * m.d4) + m.a4 * (m.b1 * m.d3 - m.b3 * m.d1));
  r.b4 = inv_det * (m.a1 * (m.b3 * m.c4 - m.b4 * m.c3) + m.a3 * (m.b4 * m.c1 - m.b1 * m.c4) + m.a4 * (m.b1 * m.c3 - m.b3 * m.c1));
  r.c1 = inv_det * (m.b1 * (m.c2 * m.d4 - m.c4 * m.d2) + m.b2 * (m.c4 * m.d1 - m.c1 * m.d4) + m.b4 * (m.c1 * m.d2 - m.c2 * m.d1));
  r.c2 = n * (m.a1 * (m.c2 * m.d4 - m.c4 * m.d2) + m.a2 * (m.c4 * m.d1 - m.c1 * m.d4) + m.a4 * (m.c1 * m.d2 - m.c2 * m.d1));
  r.c3 = inv_det * (m.a1 * (m.b2 * m.d4 - m.b4 * m.d2) + m.a2 * (m.b4 * m.d1 - m.b1 * m.d4) + m.a4 * (m.b1 * m.d2 - m.b2 * m.d1));
  r.c4 = n * (m.a1 * (m.b2 * m.c4 - m.b4 * m.c2) + m.a2 * (m.b4 * m.c1 - m.b1 * m.c4) + m.a4 * (m.b1 * m.c2 - m.b2 * m.c1));
  r.d1 = n * (m.b1 * (m.c2 * m.d3 - m.c3 * m.d2) + m.b2 * (m.c3 * m.d1 - m.c1 * m.d3) + m.b3 * (m.c1 * m.d2 - m.c2 * m.d1));
  r.d2 = inv_det * (m.a1 * (m.c2 * m.d3 - m.c3 * m.d2) + m.a2 * (m.c3 * m.d1 - m.c1 * m.d3) + m.a3 * (m.c1 * m.d2 - m.c2 * m.d1));
  r.d3 = n * (m.a1 * (m.b2 * m.d3 - m.b3 * m.d2) + m.a2 * (m.b3 * m.d1 - m.b1 * m.d3) + m.a3 * (m.b1 * m.d2 - m.b2 * m.d1));
  r.d4 = inv_det * (m.a1 * (m.b2 * m.c3 - m.b3 * m.c2) + m.a2 * (m.b3 * m.c1 - m.b1 * m.c3) + m.a3 * (m.b1 * m.c2 - m.b2 * m.c1));
  return r;
}
static inline om3 om4_to_mat3(om4 m) {
  om3 r = {m.a1, m.a2, m.a3, m.b1, m.b2, m.b3, m.c1, m.c2, m.c3};
  return r;
}
static inline float om3_det(om3 m) {
  float t11 = m.c3 * m.b2 - m.b3 * m.c2;
  float t12 = m.b3 * m.c1 - m.c3 * m.b1;
  float t13 = m.c2 * m.b1 - m.b2 * m.c1;
  return m.a1 * t11 + m.a2 * t12 + m.a3 * t13;
}
static inline ov3 om3_mul_v3(om3 m, ov3 v) {
  return ov3_new(v.x * m.a1 + v.y * m.b1 + v.z * m.c1,
                 v.x * m.a2 + v.y * m.b2 + v.z * m.c2,
                 v.x * m.a3 + v.y * m.b3 + v.z * m.c3);
}
static inline om4 om4_translate(float x, float y, float z) {
  om4 m = om4_identity(); m.d1 = x; m.d2 = y; m.d3 = z; return m;
}
static inline om4 om4_scale(float x, float y, float z) {
  om4 m = om4_identity(); m.a1 = x; m.b2 = y; m.c3 = z; return m;
}
/* the reference builds these from theta.sin_cos(); sin/cos bit patterns are an input
 * convention, not part of the hot path */
static inline om4 om4_rotate_x(float t) {
  float s = sinf(t), c = cosf(t);
  om4 m = om4_identity(); m.b2 = c; m.b3 = s; m.c2 = -s; m.c3 = c; return m;
}
static inline om4 om4_rotate_y(float t) {
  float s = sinf(t), c = cosf(t);
  om4 m = om4_identity(); m.a1 = c; m.a3 = -s; m.c1 = s; m.c3 = c; return m;
}
static inline om4 om4_rotate_z(float t) {
  float s = sinf(t), c = cosf(t);
  om4 m = om4_identity(); m.a1 = c; m.a2 = s; m.b1 = -s; m.b2 = c; return m;
}

static inline obox obox_empty(void) {
  obox b = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
  return b;
}
static inline void obox_expand_point(obox *b, ov3 p) { b->min = ov3_min(b->min, p); b->max = ov3_max(b->max, p); }
static inline void obox_expand_box(obox *b, obox o) { b->min = ov3_min(b->min, o.min); b->max = ov3_max(b->max, o.max); }
static inline ov3 obox_center(obox b) { return ov3_scale(ov3_add(b.min, b.max), 0.5f); }
static inline int obox_is_empty(obox b) { return (b.max.x < b.min.x) || (b.max.y < b.min.y) || (b.max.z < b.min.z); }
static inline float obox_surface_area(obox b) {
  float w = b.max.x - b.min.x, h = b.max.y - b.min.y, d = b.max.z - b.min.z;
  return 2.0f * (w * h + w * d + h * d);
}
/* 0 = X, 1 = Y, 2 = Z; the exact `>` cascade of box3.rs:117-133 */
static inline int obox_longest_axis(obox b) {
  float xl = b.max.x - b.min.x, yl = b.max.y - b.min.y, zl = b.max.z - b.min.z;
  if (xl > yl) { return (xl > zl) ? 0 : 2; }
  else if (yl > zl) { return 1; }
  return 2;
}
static inline obox obox_apply_matrix(obox b, om4 m) {
  if (obox_is_empty(b)) return b;
  obox r = obox_empty();
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.min.x, b.min.y, b.min.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.min.x, b.min.y, b.max.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.min.x, b.max.y, b.min.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.min.x, b.max.y, b.max.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.max.x, b.min.y, b.min.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.max.x, b.min.y, b.max.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.max.x, b.max.y, b.min.z)));
  obox_expand_point(&r, om4_mul_v3(m, ov3_new(b.max.x, b.max.y, b.max.z)));
  return r;
}
#endif
