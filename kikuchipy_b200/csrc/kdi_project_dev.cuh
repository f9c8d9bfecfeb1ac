// Device code shared by the projection kernel (kdi_project.cu) and the refinement kernel
// (kdi_refine.cu): the master-pattern handle and the per-pixel projection.
#pragma once

#include "kdi_internal.cuh"

namespace kdi_proj {
constexpr double kSqrtPiOver2 = 0.88622692545275801365;   // sqrt(pi) / 2
constexpr double kTwoOverSqrtPi = 1.1283791670955125739;  // 2 / sqrt(pi)
constexpr double kSqrtPiHalf = 1.2533141373155002512;     // sqrt(pi / 2)
}  // namespace kdi_proj

struct kdi_master_pattern {
  int mp_dtype = KDI_F32;  // storage type on the device: KDI_F32 (f32/u8/u16 sources, exact) or KDI_F64
  int npx = 0, npy = 0;    // columns, rows of the master pattern arrays
  void* upper = nullptr;
  void* lower = nullptr;
  double* dc = nullptr;  // S x 3
  // for the dictionary-generation kernel (kdi_project.cu), built on the device when the handle is created:
  // the four bilinear taps of every master-pattern position in one 16- or 32-byte element
  // [v(i, j), v(i, j+1), v(i+1, j), v(i+1, j+1)], the +1 neighbours clamped as the reference clamps them -
  // one load per pixel instead of four (the loads are gathers: the kernel was bound by the L1 lookups, one
  // 128-byte line per cycle) - and the direction cosines as three arrays (x | y | z, S doubles each)
  void* quad_upper = nullptr;
  void* quad_lower = nullptr;
  double* dc_soa = nullptr;
  int64_t S = 0;
  double scale = 0.0;
  int rescale = 0;
  double out_min = 0.0, out_max = 1.0;
};

// `P` supplies: upper, lower (const void*), npx, npy, ld (int), scale, scale_over_sqrt_pi_half (double)
// intensity of one detector pixel for one rotation (float64, as the reference computes it)
template <typename MT, typename P>
__device__ __forceinline__ double project_pixel(const P& p, const double (&m)[9], double vx,
                                                double vy, double vz) {
  // rotate_vector (_utils/numba.py:78-80) with the products precomputed per rotation.  Plain
  // IEEE multiplies and adds (no FMA contraction): for symmetric rotations the reference's terms
  // cancel EXACTLY (e.g. a rotated z of exactly 0 selects the upper hemisphere, :506); a fused
  // multiply-add would leave the rounding error of one product and could flip that choice
  const double x = __dadd_rn(__dmul_rn(m[0], vx), __dmul_rn(2.0, __dadd_rn(__dmul_rn(m[1], vz), __dmul_rn(m[2], vy))));
  const double y = __dadd_rn(__dmul_rn(m[3], vy), __dmul_rn(2.0, __dadd_rn(__dmul_rn(m[4], vx), __dmul_rn(m[5], vz))));
  const double z = __dadd_rn(__dmul_rn(m[6], vz), __dmul_rn(2.0, __dadd_rn(__dmul_rn(m[7], vy), __dmul_rn(m[8], vx))));
  // _vector2lambert (:541-566)
  // (one reciprocal instead of the reference's three divisions: at most one ulp of float64 apart;
  // an exact pole, x = y = 0, is recognised below whatever |wz| rounds to)
  const double inv = rsqrt(x * x + y * y + z * z);
  const double wx = x * inv, wy = y * inv, wz = z * inv;
  const double abs_z = fabs(wz);
  const double sqrt_z = sqrt(2.0 * (1.0 - abs_z));
  double lx = 0.0, ly = 0.0;
  if (abs_z != 1.0 && (wx != 0.0 || wy != 0.0)) {
    if (fabs(wy) <= fabs(wx)) {
      const double s = (wx > 0.0) ? 1.0 : ((wx < 0.0) ? -1.0 : 0.0);
      lx = s * sqrt_z * kdi_proj::kSqrtPiOver2;
      ly = s * sqrt_z * kdi_proj::kTwoOverSqrtPi * atan(wy / wx);
    } else {
      const double s = (wy > 0.0) ? 1.0 : ((wy < 0.0) ? -1.0 : 0.0);
      lx = s * sqrt_z * kdi_proj::kTwoOverSqrtPi * atan(wx / wy);
      ly = s * sqrt_z * kdi_proj::kSqrtPiOver2;
    }
  }
  // _get_lambert_interpolation_parameters (:638-676)
  // scale * l / sqrt(pi / 2) with the constant folded: within one ulp of float64 of the reference
  const double i_this = ly * p.scale_over_sqrt_pi_half;
  const double j_this = lx * p.scale_over_sqrt_pi_half;
  int nii = (int)(i_this + p.scale);  // truncation towards zero, like np.int32(float)
  int nij = (int)(j_this + p.scale);
  int niip = nii + 1, nijp = nij + 1;
  if (niip >= p.npx) niip = nii;
  if (nijp >= p.npy) nijp = nij;
  if (nii < 0) nii = niip;
  if (nij < 0) nij = nijp;
  const double di = i_this - (double)nii + p.scale;
  const double dj = j_this - (double)nij + p.scale;
  const double dim = 1.0 - di, djm = 1.0 - dj;
  // _get_pixel_from_master_pattern (:703-708), hemisphere by the sign of the ROTATED z (:506)
  const MT* mp = reinterpret_cast<const MT*>(z >= 0.0 ? p.upper : p.lower);
  const double v00 = (double)__ldg(mp + (int64_t)nii * p.ld + nij);
  const double v10 = (double)__ldg(mp + (int64_t)niip * p.ld + nij);
  const double v01 = (double)__ldg(mp + (int64_t)nii * p.ld + nijp);
  const double v11 = (double)__ldg(mp + (int64_t)niip * p.ld + nijp);
  return v00 * dim * djm + v10 * di * djm + v01 * dim * dj + v11 * di * dj;
}

// ---- the same pixel with the library calls replaced (dictionary generation, kdi_project.cu) ----------
// `project_pixel` spends two thirds of its ~340 instructions in the CUDA math library: atan() twice (the
// two branches of the Lambert projection diverge inside a warp), a division in each, sqrt(), rsqrt(), with
// their special-case paths.  Here: one division for both branches, seeded by the hardware's 20-bit
// reciprocal and finished by a cubic Newton step + one residual correction; the same for the square root;
// an odd polynomial of 11 terms (near-minimax) for atan on |r| <= tan(pi/8) after the reduction
// atan(a/b) = pi/4 + atan((a - b)/(a + b)), folded into the one division; the normalisation of the rotated
// vector as a Taylor step (it is a unit vector rotated by a unit quaternion: |n^2 - 1| ~ 1e-16); bilinear
// blend as three lerps over ONE 16-byte load of the pixel's four taps (tap tables, kdi_master_pattern).  Every intermediate stays within a few ulp of float64 of the reference's value
// (the reference itself is compiled with fastmath and is no better defined), and the float32 results
// agree with the goldens exactly as often as the library version's do (tests/test_gpu_projection.py).
// `m2`: rotation products with the factor 2 folded into the off-diagonal terms (exact).
// `P` supplies in addition: quad_upper, quad_lower (the tap tables of kdi_master_pattern).
namespace kdi_proj {
__device__ __forceinline__ double rcp_seed(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  return y;
}
__device__ __forceinline__ double rsqrt_seed(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  return y;
}
// 1 / sqrt(t) to float64 rounding for any normal t > 0: hardware seed (20 bits), one cubic step, one
// residual correction (the library's 1.0 / sqrt(t) costs three times the instructions)
__device__ __forceinline__ double rsqrt_full(double t) {
  double y = rsqrt_seed(t);
  const double e = fma(-(t * y), y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
  const double e2 = fma(-(t * y), y, 1.0);
  return fma(0.5 * y, e2, y);
}
constexpr double kTanPiOver8 = 0.41421356237309504880;
constexpr double kPiOver4 = 0.78539816339744830962;
// atan(r) / r as a polynomial in u = r^2 on [0, tan^2(pi/8)]: relative error 2.2e-16 (interpolation at
// Chebyshev nodes, computed with 60 digits; tools/probes/atan_fit.py).  In constant memory: the
// multiply-adds take their coefficient operand straight from the constant bank (as immediates each one
// costs two moves per pixel, and the loop is issue-bound)
__constant__ double kAtanC[11] = {0x1.0000000000000p+0,  -0x1.55555555551e4p-2, 0x1.9999999934e50p-3,  -0x1.2492490095732p-3,
                                  0x1.c71c648aca6a2p-4,  -0x1.745baa2843f92p-4, 0x1.3afb01bbe7601p-4,  -0x1.0ffbe63b09941p-4,
                                  0x1.d1fd545f28ae3p-5,  -0x1.64314395b2626p-5, 0x1.5a482e20112ecp-6};
__device__ __forceinline__ double atan_over_r(double u) {
  double p = kAtanC[10];
#pragma unroll
  for (int k = 9; k >= 0; --k) p = fma(p, u, kAtanC[k]);
  return p;
}
// |x| by clearing the sign bit (an integer instruction; fabs() of a value that goes through a select is
// materialised with a double-precision add)
__device__ __forceinline__ double abs_bits(double x) {
  return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}
}  // namespace kdi_proj

// In two halves, so that a caller can keep the tap load of one pixel in flight while it computes the
// coordinates of the next (the loads are L2 gathers, several hundred cycles each): `..._fetch` ends with
// the load issued, `..._blend` consumes it.
template <typename MT>
struct kdi_lean_taps {
  double di, dj;
  MT v00, v10, v01, v11;
};

template <typename MT, typename P>
__device__ __forceinline__ kdi_lean_taps<MT> project_pixel_lean_fetch(const P& p, const double (&m2)[9], double vx,
                                                                     double vy, double vz) {
  using namespace kdi_proj;
  // rotate_vector: z with separately rounded products and sums (its sign picks the hemisphere and
  // must be exactly 0 where the reference's terms cancel); x and y may fuse
  const double z = __dadd_rn(__dmul_rn(m2[6], vz), __dadd_rn(__dmul_rn(m2[7], vy), __dmul_rn(m2[8], vx)));
  const double x = fma(m2[0], vx, fma(m2[1], vz, m2[2] * vy));
  const double y = fma(m2[3], vy, fma(m2[4], vx, m2[5] * vz));
  // 1 / |v|
  const double n2 = fma(x, x, fma(y, y, z * z));
  const double dn = n2 - 1.0;
  double inv = fma(dn, fma(dn, 0.375, -0.5), 1.0);  // (1 + d)^(-1/2) to 5/16 d^3
  if (!(fabs(dn) < 1e-5)) inv = rsqrt(n2);          // not a unit quaternion / direction: the long way
  const double wx = x * inv, wy = y * inv, wz = z * inv;
  // _vector2lambert
  // (the Taylor step may leave |z| one ulp above 1: clamped on the bit pattern, an integer compare)
  const int hz = __double2hiint(wz) & 0x7fffffff;
  const double abs_z = hz >= 0x3ff00000 ? 1.0 : __hiloint2double(hz, __double2loint(wz));
  const double t = fma(-2.0, abs_z, 2.0);    // 2 (1 - |z|), same rounding
  double ys = rsqrt_seed(t);
  {
    const double ty = t * ys, e = fma(-ty, ys, 1.0);
    ys = fma(ys * e, fma(0.375, e, 0.5), ys);
  }
  const double sqrt_z = t * ys;
  const bool first = fabs(wy) <= fabs(wx);
  const double den = first ? wx : wy, num = first ? wy : wx;
  const double a = abs_bits(num), b = abs_bits(den);  // 0 <= a <= b
  const bool red = a > kTanPiOver8 * b;
  const double rn = red ? a - b : a, rd = red ? a + b : b;
  double yr = rcp_seed(rd);
  {
    const double e = fma(-rd, yr, 1.0);
    yr = fma(yr, fma(e, e, e), yr);
  }
  double r = rn * yr;
  r = fma(fma(-rd, r, rn), yr, r);
  const double u = r * r;
  const double q = fma(r, atan_over_r(u), red ? kPiOver4 : 0.0);  // atan(a / b) in [0, pi/4]
  const double major = copysign(sqrt_z * kSqrtPiOver2, den);
  const double minor = copysign((sqrt_z * kTwoOverSqrtPi) * q, num);
  const bool pole = (abs_z == 1.0) || (b == 0.0);
  const double lx = pole ? 0.0 : (first ? major : minor);
  const double ly = pole ? 0.0 : (first ? minor : major);
  // _get_lambert_interpolation_parameters
  const double i_this = ly * p.scale_over_sqrt_pi_half;
  const double j_this = lx * p.scale_over_sqrt_pi_half;
  // (truncation towards zero never goes negative here, so the reference's `nii < 0` branch is dead; the
  // clamp of the +1 neighbours lives in the tap table; the min() only guards the table against a caller's
  // scale that does not match the master pattern)
  const int nii = min(__double2int_rz(i_this + p.scale), p.npy - 1);
  const int nij = min(__double2int_rz(j_this + p.scale), p.npx - 1);
  const double di = i_this - (double)nii + p.scale;
  const double dj = j_this - (double)nij + p.scale;
  // _get_pixel_from_master_pattern: the four taps in one load (hemisphere by the sign of the ROTATED z)
  // (element offsets fit 32 bits: the master pattern has at most 32768 x 32768 values)
  const int at = nii * p.ld + nij;
  kdi_lean_taps<MT> t4;
  t4.di = di;
  t4.dj = dj;
  if constexpr (sizeof(MT) == 4) {
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(z >= 0.0 ? p.quad_upper : p.quad_lower) + at);
    t4.v00 = q4.x; t4.v01 = q4.y; t4.v10 = q4.z; t4.v11 = q4.w;
  } else {
    const double2* q2 = reinterpret_cast<const double2*>(z >= 0.0 ? p.quad_upper : p.quad_lower) + 2 * (int64_t)at;
    const double2 lo = __ldg(q2), hi = __ldg(q2 + 1);
    t4.v00 = lo.x; t4.v01 = lo.y; t4.v10 = hi.x; t4.v11 = hi.y;
  }
  return t4;
}

template <typename MT>
__device__ __forceinline__ double project_pixel_lean_blend(const kdi_lean_taps<MT>& t4) {
  const double v00 = (double)t4.v00, v10 = (double)t4.v10, v01 = (double)t4.v01, v11 = (double)t4.v11;
  const double a0 = fma(t4.di, v10 - v00, v00), a1 = fma(t4.di, v11 - v01, v01);
  return fma(t4.dj, a1 - a0, a0);
}

template <typename MT, typename P>
__device__ __forceinline__ double project_pixel_lean(const P& p, const double (&m2)[9], double vx,
                                                     double vy, double vz) {
  return project_pixel_lean_blend<MT>(project_pixel_lean_fetch<MT>(p, m2, vx, vy, vz));
}
