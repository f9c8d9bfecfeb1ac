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
