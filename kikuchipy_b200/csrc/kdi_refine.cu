// Refinement of orientations and/or projection centres on the device.
//
// Replaces, for the reference's default optimiser (scipy.optimize.minimize, Nelder-Mead),
// /root/reference/src/kikuchipy/indexing/_refinement/
//   _solvers.py:50-73    _prepare_pattern (cast, optional rescale to [-1, 1], centre, squared norm)
//   _solvers.py:79-254   _refine_orientation_solver_scipy (all starts of a pattern, best score wins)
//   _solvers.py:257-345  _refine_pc_solver_scipy, :348-470 _refine_orientation_pc_solver_scipy
//   _objective_functions.py:36-190  the three objective functions (1 - NCC of the pattern and a
//                        pattern projected from the master pattern)
// and what they call: _utils/numba.py:44-58 rotation_from_euler, _utils/_gnonomic_bounds.py:23-62,
// signals/util/_master_pattern.py:133-204 (direction cosines of a projection centre), :449-527
// (single-pattern projection), similarity_metrics/_normalized_cross_correlation.py:200-225
// (_ncc_single_patterns_1d_float32_exp_centered: float32 arithmetic, widened to float64).
// The simplex search restates SciPy's _minimize_neldermead (third party; scipy/optimize/
// _optimize.py): simplex construction (5 % / 0.00025 steps), bound handling (clip, reflect the
// initial simplex into the interior), the (1 + rho) * xbar - rho * worst coefficient arithmetic
// with every product and sum rounded separately (no FMA contraction - the search compares
// objective values, so a different rounding of a trial point changes the trajectory), the
// termination test and the maxfev abort semantics.
//
// One CTA per experimental pattern.  The centred pattern lives in shared memory for the whole
// search; each objective evaluation projects the kept detector pixels block-wide (float64
// geometry, like the reference), stages the simulated pattern as float32 in shared memory and
// reduces the three sums.  Every thread runs the (tiny) simplex bookkeeping redundantly on the
// broadcast objective value, so the control flow is uniform and needs no extra barriers.
// Bound: fp64 pipe (about 300 double operations per pixel and evaluation).
#include <cfloat>
#include <climits>
#include <cstdlib>

#include "kdi_internal.cuh"
#include "kdi_project_dev.cuh"

namespace {

constexpr int kRefThreads = 256;
constexpr int kRefWarps = kRefThreads / 32;

struct RefineParams {
  // what project_pixel_lean needs
  const void* upper;
  const void* lower;
  const void* quad_upper;
  const void* quad_lower;
  int npx, npy, ld;
  double scale, scale_over_sqrt_pi_half;
  // experimental patterns
  const void* pat;
  const int64_t* pat_rows;  // objective mode: source pattern of each row (null = row i is pattern i)
  int pat_dtype;
  int64_t S;            // detector pixels per pattern (nrows * ncols)
  const int32_t* cols;  // kept pixels (signal mask) or null
  int64_t s_eff;
  int rescale;
  int64_t n_patterns;
  int n_starts;
  const double* x0;     // n_patterns x n_starts x NV
  const double* lb;     // same shape, or null
  const double* ub;
  const double* quat;   // PC mode: n_patterns x 4
  const double* pcs;    // ORI mode with one PC per pattern: n_patterns x 3 (null: fixed direction cosines)
  const double* dc;     // S x 3 direction cosines of the whole detector (ORI mode, fixed PC)
  int nrows, ncols;
  double om[9];         // detector -> sample, row-major
  double xatol, fatol;
  long long maxiter, maxfev;
  int adaptive;
  double* out;          // n_patterns x out_stride
  int out_stride;
  // patterns too large for shared memory (more than 25 600 matched pixels): per-CTA staging in global
  // memory instead (2 * pitch floats per CTA, L2-resident; the kernel is bound by the float64 geometry,
  // not by these reads).  null = shared memory
  float* gws;
};

__device__ __forceinline__ double block_sum(double v, double (*red)[kRefWarps]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[0][warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kRefWarps; ++w) t += red[0][w];
  return t;
}

__device__ __forceinline__ void block_sum2(double& a, double& b, double (*red)[kRefWarps]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { red[0][warp] = a; red[1][warp] = b; }
  __syncthreads();
  double ta = 0.0, tb = 0.0;
#pragma unroll
  for (int w = 0; w < kRefWarps; ++w) { ta += red[0][w]; tb += red[1][w]; }
  a = ta;
  b = tb;
}

__device__ __forceinline__ void block_minmax(float& lo, float& hi, double (*red)[kRefWarps]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { red[0][warp] = (double)lo; red[1][warp] = (double)hi; }
  __syncthreads();
  lo = (float)red[0][0];
  hi = (float)red[1][0];
#pragma unroll
  for (int w = 1; w < kRefWarps; ++w) { lo = fminf(lo, (float)red[0][w]); hi = fmaxf(hi, (float)red[1][w]); }
}

__device__ __forceinline__ float load_pixel(const void* base, int dtype, int64_t i) {
  switch (dtype) {
    case KDI_U8: return (float)reinterpret_cast<const uint8_t*>(base)[i];
    case KDI_U16: return (float)reinterpret_cast<const uint16_t*>(base)[i];
    case KDI_F32: return reinterpret_cast<const float*>(base)[i];
    default: return (float)reinterpret_cast<const double*>(base)[i];
  }
}

// 1 - NCC(pattern, projection) for the control variables x.  MODE 0: x = Euler angles (direction
// cosines fixed, or from the pattern's own PC); 1: x = PC, rotation `quat`; 2: x = Euler + PC.
template <int MODE>
__device__ double evaluate(const RefineParams& p, const double* x, const double* quat, const double* pc_fixed,
                           const float* e, float* v, float sqnorm, double (*red)[kRefWarps]) {
  double a, b, c, d;
  if (MODE == 1) {
    a = quat[0]; b = quat[1]; c = quat[2]; d = quat[3];
  } else {  // rotation_from_euler (_utils/numba.py:44-58)
    const double sigma = 0.5 * (x[0] + x[2]), delta = 0.5 * (x[0] - x[2]);
    double sb, cb, ss, cs, sd, cd;
    sincos(0.5 * x[1], &sb, &cb);
    sincos(sigma, &ss, &cs);
    sincos(delta, &sd, &cd);
    a = cb * cs; b = -sb * cd; c = -sb * sd; d = -cb * ss;
    if (a < 0.0) { a = -a; b = -b; c = -c; d = -d; }
  }
  const double aa = __dmul_rn(a, a), bb = __dmul_rn(b, b), cc = __dmul_rn(c, c), dd = __dmul_rn(d, d);
  const double ac = __dmul_rn(a, c), ab = __dmul_rn(a, b), ad = __dmul_rn(a, d);
  const double bc = __dmul_rn(b, c), bd = __dmul_rn(b, d), cd2 = __dmul_rn(c, d);
  const double m[9] = {__dadd_rn(__dadd_rn(__dadd_rn(aa, bb), -cc), -dd), __dadd_rn(ac, bd), __dadd_rn(bc, -ad),
                       __dadd_rn(__dadd_rn(__dadd_rn(aa, -bb), cc), -dd), __dadd_rn(ad, bc), __dadd_rn(cd2, -ab),
                       __dadd_rn(__dadd_rn(__dadd_rn(aa, -bb), -cc), dd), __dadd_rn(ab, cd2), __dadd_rn(bd, -ac)};
  // (the factor 2 of the off-diagonal terms folded in: exact)
  const double m2[9] = {m[0], 2.0 * m[1], 2.0 * m[2], m[3], 2.0 * m[4], 2.0 * m[5], m[6], 2.0 * m[7], 2.0 * m[8]};
  // direction cosines from a projection centre (get_gnomonic_bounds + _get_direction_cosines_for_fixed_pc)
  const double* pc = (MODE == 1) ? x : (MODE == 2 ? x + 3 : pc_fixed);
  double gx0 = 0, gy0 = 0, xs = 0, ys = 0, xh = 0, yh = 0, pcz = 0;
  if (pc) {
    const double aspect = (double)p.ncols / (double)p.nrows;
    pcz = pc[2];
    const double x_min = -aspect * (pc[0] / pcz), x_max = aspect * (1.0 - pc[0]) / pcz;
    const double y_min = -(1.0 - pc[1]) / pcz, y_max = pc[1] / pcz;
    xs = (x_max - x_min) / (double)p.ncols;
    ys = (y_max - y_min) / (double)p.nrows;
    gx0 = x_min;
    gy0 = y_max;
    xh = xs / 2.0;
    yh = ys / 2.0;
  }
  double sum = 0.0;
  for (int64_t j = threadIdx.x; j < p.s_eff; j += kRefThreads) {
    const int64_t idx = p.cols ? (int64_t)p.cols[j] : j;
    double vx, vy, vz;
    if (pc) {
      // (32-bit: a 64-bit division is a ~40-instruction routine, and this loop is issue-bound)
      const unsigned r = (unsigned)idx / (unsigned)p.ncols, cidx = (unsigned)idx - r * (unsigned)p.ncols;
      const double gx = (gx0 + (double)cidx * xs + xh) * pcz;
      const double gy = (gy0 + (double)r * (-ys) - yh) * pcz;
      vx = gx * p.om[0] + gy * p.om[1] + pcz * p.om[2];
      vy = gx * p.om[3] + gy * p.om[4] + pcz * p.om[5];
      vz = gx * p.om[6] + gy * p.om[7] + pcz * p.om[8];
      const double inv = kdi_proj::rsqrt_full(vx * vx + vy * vy + vz * vz);
      vx *= inv; vy *= inv; vz *= inv;
    } else {
      vx = __ldg(p.dc + 3 * idx); vy = __ldg(p.dc + 3 * idx + 1); vz = __ldg(p.dc + 3 * idx + 2);
    }
    const float fv = (float)project_pixel_lean<float>(p, m2, vx, vy, vz);
    v[j] = fv;
    sum += (double)fv;
  }
  sum = block_sum(sum, red);
  const float mean = (float)(sum / (double)p.s_eff);
  double s1 = 0.0, s2 = 0.0;
  for (int64_t j = threadIdx.x; j < p.s_eff; j += kRefThreads) {
    const float cv = __fsub_rn(v[j], mean);
    s1 += (double)__fmul_rn(e[j], cv);
    s2 += (double)__fmul_rn(cv, cv);
  }
  block_sum2(s1, s2, red);
  const float ncc = __fdiv_rn((float)s1, __fsqrt_rn(__fmul_rn(sqnorm, (float)s2)));
  return 1.0 - (double)ncc;
}

__device__ __forceinline__ bool f_less(double a, double b) { return (a < b) || (b != b && a == a); }
__device__ __forceinline__ double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

// _prepare_pattern (_solvers.py:50-73): the kept pixels of source pattern `src_row` centred in e[],
// optionally rescaled to [-1, 1] first; returns the squared norm
__device__ __forceinline__ float prepare_pattern(const RefineParams& p, int64_t src_row, float* e,
                                                 double (*red)[kRefWarps]) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int64_t j = threadIdx.x; j < p.s_eff; j += kRefThreads) {
    const float val = load_pixel(p.pat, p.pat_dtype, src_row * p.S + (p.cols ? (int64_t)p.cols[j] : j));
    e[j] = val;
    lo = fminf(lo, val);
    hi = fmaxf(hi, val);
  }
  if (p.rescale) {  // (pattern - min) / float(max - min) * 2 - 1, the quotient in float64
    block_minmax(lo, hi, red);
    const double range = (double)__fsub_rn(hi, lo);
    for (int64_t j = threadIdx.x; j < p.s_eff; j += kRefThreads)
      e[j] = (float)((double)__fsub_rn(e[j], lo) / range * 2.0 + (-1.0));
  }
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < p.s_eff; j += kRefThreads) s += (double)e[j];
  s = block_sum(s, red);
  const float mean = (float)(s / (double)p.s_eff);
  double sq = 0.0;
  for (int64_t j = threadIdx.x; j < p.s_eff; j += kRefThreads) {
    const float cv = __fsub_rn(e[j], mean);
    e[j] = cv;
    sq += (double)__fmul_rn(cv, cv);
  }
  sq = block_sum(sq, red);
  return (float)sq;
}

// Objective values only: row i = 1 - NCC of pattern pat_rows[i] and the projection for each of its
// n_starts parameter sets x0[i][k] (with quat[i][k] in PC mode, pcs[i] in orientation mode).  This is the
// function the reference hands to scipy.optimize (_objective_functions.py:36-190); optimisers other than
// Nelder-Mead run on the host and call it in batches (kikuchipy_b200/refinement.py).
template <int MODE, int NV>
__global__ void __launch_bounds__(kRefThreads) kdi_refine_objective_kernel(const RefineParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t pitch = (p.s_eff + 3) & ~(int64_t)3;
  float* e = p.gws ? p.gws + (int64_t)blockIdx.x * 2 * pitch : reinterpret_cast<float*>(smem_raw);
  float* v = e + pitch;
  __shared__ double red[2][kRefWarps];
  for (int64_t row = blockIdx.x; row < p.n_patterns; row += gridDim.x) {
    __syncthreads();
    const float sqnorm = prepare_pattern(p, p.pat_rows ? p.pat_rows[row] : row, e, red);
    for (int st = 0; st < p.n_starts; ++st) {
      const int64_t so = (row * p.n_starts + st) * NV;
      const double* quat = (MODE == 1) ? p.quat + (row * p.n_starts + st) * 4 : nullptr;
      const double* pcf = (MODE == 0 && p.pcs) ? p.pcs + row * 3 : nullptr;
      const double f = evaluate<MODE>(p, p.x0 + so, quat, pcf, e, v, sqnorm, red);
      if (threadIdx.x == 0) p.out[row * p.n_starts + st] = f;
    }
  }
}

template <int MODE, int NV, int MINB>
__global__ void __launch_bounds__(kRefThreads, MINB) kdi_refine_kernel(const RefineParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int64_t pitch = (p.s_eff + 3) & ~(int64_t)3;
  float* e = p.gws ? p.gws + (int64_t)blockIdx.x * 2 * pitch : reinterpret_cast<float*>(smem_raw);
  float* v = e + pitch;
  __shared__ double red[2][kRefWarps];

  for (int64_t row = blockIdx.x; row < p.n_patterns; row += gridDim.x) {
    __syncthreads();
    const float sqnorm = prepare_pattern(p, row, e, red);

    // ---- Nelder-Mead from every start; the best score wins (_solvers.py:236-254) ----
    double rho = 1.0, chi = 2.0, psi = 0.5, sigma = 0.5;
    if (p.adaptive) {
      const double dim = (double)NV;
      chi = 1.0 + 2.0 / dim;
      psi = 0.75 - 1.0 / (2.0 * dim);
      sigma = 1.0 - 1.0 / dim;
    }
    double best_ncc = 0.0, best_x[NV];
    long long best_nfev = 0;
    int best_start = -1;
    for (int st = 0; st < p.n_starts; ++st) {
      const int64_t so = (row * p.n_starts + st) * NV;
      const double* quat = (MODE == 1) ? p.quat + (row * p.n_starts + st) * 4 : nullptr;
      const double* pcf = (MODE == 0 && p.pcs) ? p.pcs + row * 3 : nullptr;
      const bool bounded = p.lb != nullptr;
      double lbv[NV], ubv[NV];
      double sim[NV + 1][NV], fs[NV + 1];
      for (int i = 0; i < NV; ++i) {
        lbv[i] = bounded ? p.lb[so + i] : 0.0;
        ubv[i] = bounded ? p.ub[so + i] : 0.0;
        const double xi = p.x0[so + i];
        sim[0][i] = bounded ? clipd(xi, lbv[i], ubv[i]) : xi;
      }
      for (int k = 0; k < NV; ++k) {
        for (int i = 0; i < NV; ++i) sim[k + 1][i] = sim[0][i];
        sim[k + 1][k] = (sim[0][k] != 0.0) ? __dmul_rn(1.0 + 0.05, sim[0][k]) : 0.00025;
      }
      if (bounded)
        for (int k = 0; k <= NV; ++k)
          for (int i = 0; i < NV; ++i) {
            double t = sim[k][i];
            if (t > ubv[i]) t = __dsub_rn(__dmul_rn(2.0, ubv[i]), t);
            sim[k][i] = clipd(t, lbv[i], ubv[i]);
          }
      long long fcalls = 0;
      // the wrapped objective of SciPy: refuses (raises) once maxfev calls have been made
      auto feval = [&](const double* xx, double& f) -> bool {
        if (fcalls >= p.maxfev) return false;
        ++fcalls;
        f = evaluate<MODE>(p, xx, quat, pcf, e, v, sqnorm, red);
        return true;
      };
      auto sort_simplex = [&]() {  // stable, ascending, NaN last (np.argsort of a handful of values)
        for (int i = 1; i <= NV; ++i) {
          const double fi = fs[i];
          double xi[NV];
          for (int t = 0; t < NV; ++t) xi[t] = sim[i][t];
          int j = i - 1;
          while (j >= 0 && f_less(fi, fs[j])) {
            fs[j + 1] = fs[j];
            for (int t = 0; t < NV; ++t) sim[j + 1][t] = sim[j][t];
            --j;
          }
          fs[j + 1] = fi;
          for (int t = 0; t < NV; ++t) sim[j + 1][t] = xi[t];
        }
      };
      for (int k = 0; k <= NV; ++k) fs[k] = INFINITY;
      for (int k = 0; k <= NV; ++k)
        if (!feval(sim[k], fs[k])) break;
      sort_simplex();
      long long iterations = 1;
      while (fcalls < p.maxfev && iterations < p.maxiter) {
        double dx = 0.0, df = 0.0;
        for (int k = 1; k <= NV; ++k) {
          for (int i = 0; i < NV; ++i) dx = fmax(dx, fabs(__dsub_rn(sim[k][i], sim[0][i])));
          df = fmax(df, fabs(__dsub_rn(fs[0], fs[k])));
        }
        if (dx <= p.xatol && df <= p.fatol) break;
        do {  // one simplex step; `break` = SciPy's _MaxFuncCallError leaving the try block
          double xbar[NV], xt[NV], ft;
          for (int i = 0; i < NV; ++i) {
            double t = sim[0][i];
            for (int k = 1; k < NV; ++k) t = __dadd_rn(t, sim[k][i]);
            xbar[i] = __ddiv_rn(t, (double)NV);
          }
          auto combine = [&](double c1, double c2, bool plus) {
            for (int i = 0; i < NV; ++i) {
              const double t1 = __dmul_rn(c1, xbar[i]), t2 = __dmul_rn(c2, sim[NV][i]);
              const double t = plus ? __dadd_rn(t1, t2) : __dsub_rn(t1, t2);
              xt[i] = bounded ? clipd(t, lbv[i], ubv[i]) : t;
            }
          };
          auto accept = [&](const double* xx, double f) {
            for (int i = 0; i < NV; ++i) sim[NV][i] = xx[i];
            fs[NV] = f;
          };
          double xr[NV], fxr;
          combine(1.0 + rho, rho, false);
          for (int i = 0; i < NV; ++i) xr[i] = xt[i];
          if (!feval(xr, fxr)) break;
          bool shrink = false;
          if (fxr < fs[0]) {
            combine(1.0 + rho * chi, rho * chi, false);
            if (!feval(xt, ft)) break;
            if (ft < fxr) accept(xt, ft);
            else accept(xr, fxr);
          } else if (fxr < fs[NV - 1]) {
            accept(xr, fxr);
          } else if (fxr < fs[NV]) {
            combine(1.0 + psi * rho, psi * rho, false);
            if (!feval(xt, ft)) break;
            if (ft <= fxr) accept(xt, ft);
            else shrink = true;
          } else {
            combine(1.0 - psi, psi, true);
            if (!feval(xt, ft)) break;
            if (ft < fs[NV]) accept(xt, ft);
            else shrink = true;
          }
          if (shrink) {
            bool aborted = false;
            for (int k = 1; k <= NV && !aborted; ++k) {
              for (int i = 0; i < NV; ++i) {
                const double t = __dadd_rn(sim[0][i], __dmul_rn(sigma, __dsub_rn(sim[k][i], sim[0][i])));
                sim[k][i] = bounded ? clipd(t, lbv[i], ubv[i]) : t;
              }
              aborted = !feval(sim[k], fs[k]);
            }
            if (aborted) break;
          }
          ++iterations;
        } while (false);
        sort_simplex();
      }
      double fmin_v = fs[0];  // np.min(fsim)
      for (int k = 1; k <= NV; ++k) fmin_v = (fs[k] < fmin_v || fs[k] != fs[k]) ? fs[k] : fmin_v;
      const double ncc = 1.0 - fmin_v;
      if (best_start < 0 || ncc > best_ncc) {  // np.argmax: the first maximum
        best_ncc = ncc;
        best_nfev = fcalls;
        best_start = st;
        for (int i = 0; i < NV; ++i) best_x[i] = sim[0][i];
      }
    }
    if (threadIdx.x == 0) {
      double* o = p.out + row * p.out_stride;
      o[0] = best_ncc;
      o[1] = (double)best_nfev;
      for (int i = 0; i < NV; ++i) o[2 + i] = best_x[i];
      if (p.n_starts > 1) o[2 + NV] = (double)best_start;
    }
  }
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// objective == true: only evaluate the objective at the given points (pattern_rows: source pattern per row)
static int refine_impl(kdi_ctx* ctx, const kdi_master_pattern* mp, int mode, const void* patterns,
                       int pat_loc, int pat_dtype, int64_t n_patterns, int nrows, int ncols, int rescale,
                       const double* x0, int n_starts, const double* lower, const double* upper,
                       const double* rotations, const double* pcs, const double* om_detector_to_sample,
                       const kdi_refine_options* opt, double* results_out, bool objective,
                       const int64_t* pattern_rows, int64_t n_source_patterns) {
  if (!ctx) return KDI_EINVAL;
  if (!mp || !patterns || !x0 || (!opt && !objective) || !results_out) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: NULL argument");
  static const kdi_refine_options no_options = {0.0, 0.0, -1, -1, 0};
  if (objective) opt = &no_options;
  if (mode < KDI_REFINE_ORI || mode > KDI_REFINE_ORI_PC) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: unknown mode %d", mode);
  if (n_patterns < 0 || n_starts < 1 || nrows < 1 || ncols < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: bad shape");
  if (!kdi_dtype_size(pat_dtype)) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: unknown pattern dtype %d", pat_dtype);
  if ((lower == nullptr) != (upper == nullptr)) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: give both bounds or none");
  if (mode == KDI_REFINE_PC && (!rotations || (n_starts != 1 && !objective)))
    return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: PC refinement needs one rotation per pattern and one start");
  if (mode == KDI_REFINE_ORI && pcs == nullptr && (int64_t)nrows * ncols != mp->S)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: detector has %lld pixels, the master pattern's direction cosines %lld",
                    (long long)nrows * ncols, (long long)mp->S);
  if ((mode != KDI_REFINE_ORI || pcs) && !om_detector_to_sample)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: the detector-to-sample matrix is needed when projection centres vary");
  if (mp->mp_dtype != KDI_F32)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_refine: the master pattern must be float32 (the reference converts it, _refinement.py:1313-1319)");
  if (!(opt->xatol >= 0.0) || !(opt->fatol >= 0.0)) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: negative tolerance");
  if (n_patterns == 0) return KDI_OK;
  const int64_t S = (int64_t)nrows * ncols;
  if (ctx->mask_S && ctx->mask_S != S)
    return kdi_fail(ctx, KDI_EINVAL, "signal mask has %lld pixels, the detector %lld", (long long)ctx->mask_S, (long long)S);
  const int nv = mode == KDI_REFINE_ORI_PC ? 6 : 3;
  const int64_t s_eff = ctx->mask_S ? ctx->mask_kept : S;
  size_t smem = 2 * (size_t)((s_eff + 3) & ~(int64_t)3) * sizeof(float);
  if (s_eff < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_refine: the signal mask excludes every pixel");
  // large patterns (e.g. 240 x 240, 480 x 480): a bounded resident grid with its staging in global memory
  const bool global_staging = smem > 200 * 1024;
  const int64_t max_ctas = (int64_t)ctx->sm_count * 4;
  const size_t gws_bytes = global_staging ? (size_t)std::min<int64_t>(n_patterns, max_ctas) * smem : 0;
  if (global_staging) smem = 0;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;

  // workspace: [patterns if on the host] [x0] [lb] [ub] [quat] [pcs] [out]
  const size_t esz = kdi_dtype_size(pat_dtype);
  const size_t n_x = (size_t)n_patterns * n_starts * nv;
  const int out_stride = objective ? n_starts : 2 + nv + (n_starts > 1 ? 1 : 0);
  // (objective mode: `patterns` holds n_source_patterns patterns, the rows name theirs)
  const int64_t n_stored = (objective && pattern_rows) ? n_source_patterns : n_patterns;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = up256(o + bytes); return at; };
  const size_t o_pat = pat_loc == KDI_HOST ? take((size_t)n_stored * S * esz) : 0;
  const size_t o_rows = (objective && pattern_rows) ? take((size_t)n_patterns * 8) : 0;
  const size_t o_x0 = take(n_x * 8), o_lb = lower ? take(n_x * 8) : 0, o_ub = lower ? take(n_x * 8) : 0;
  const size_t o_q = rotations ? take((size_t)n_patterns * n_starts * 32) : 0;
  const size_t o_pc = pcs ? take((size_t)n_patterns * 24) : 0;
  const size_t o_out = take((size_t)n_patterns * out_stride * 8);
  const size_t o_gws = global_staging ? take(gws_bytes) : 0;
  KDI_TRY(kdi_ws2_reserve(ctx, o));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  ctx->tm = kdi_timings();
  if (pat_loc == KDI_HOST) {
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_pat, patterns, (size_t)n_stored * S * esz, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += n_stored * S * (int64_t)esz;
  }
  if (objective && pattern_rows) {
    for (int64_t i = 0; i < n_patterns; ++i)
      if (pattern_rows[i] < 0 || pattern_rows[i] >= n_source_patterns)
        return kdi_fail(ctx, KDI_EINVAL, "kdi_refine_objective: pattern row %lld out of range", (long long)pattern_rows[i]);
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_rows, pattern_rows, (size_t)n_patterns * 8, cudaMemcpyHostToDevice, st));
  }
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_x0, x0, n_x * 8, cudaMemcpyHostToDevice, st));
  if (lower) {
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_lb, lower, n_x * 8, cudaMemcpyHostToDevice, st));
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_ub, upper, n_x * 8, cudaMemcpyHostToDevice, st));
  }
  if (rotations) KDI_CUDA(ctx, cudaMemcpyAsync(w + o_q, rotations, (size_t)n_patterns * n_starts * 32, cudaMemcpyHostToDevice, st));
  if (pcs) KDI_CUDA(ctx, cudaMemcpyAsync(w + o_pc, pcs, (size_t)n_patterns * 24, cudaMemcpyHostToDevice, st));

  RefineParams p = {};
  p.upper = mp->upper;
  p.lower = mp->lower;
  p.quad_upper = mp->quad_upper;
  p.quad_lower = mp->quad_lower;
  p.npx = mp->npx;
  p.npy = mp->npy;
  p.ld = mp->npx;
  p.scale = mp->scale;
  p.scale_over_sqrt_pi_half = mp->scale / kdi_proj::kSqrtPiHalf;
  p.pat = pat_loc == KDI_HOST ? (const void*)(w + o_pat) : patterns;
  p.pat_rows = (objective && pattern_rows) ? reinterpret_cast<const int64_t*>(w + o_rows) : nullptr;
  p.pat_dtype = pat_dtype;
  p.S = S;
  p.cols = ctx->mask_S ? ctx->d_cols : nullptr;
  p.s_eff = s_eff;
  p.rescale = rescale != 0;
  p.n_patterns = n_patterns;
  p.n_starts = n_starts;
  p.x0 = reinterpret_cast<const double*>(w + o_x0);
  p.lb = lower ? reinterpret_cast<const double*>(w + o_lb) : nullptr;
  p.ub = lower ? reinterpret_cast<const double*>(w + o_ub) : nullptr;
  p.quat = rotations ? reinterpret_cast<const double*>(w + o_q) : nullptr;
  p.pcs = pcs ? reinterpret_cast<const double*>(w + o_pc) : nullptr;
  p.dc = mp->dc;
  p.nrows = nrows;
  p.ncols = ncols;
  for (int i = 0; i < 9; ++i) p.om[i] = om_detector_to_sample ? om_detector_to_sample[i] : 0.0;
  p.xatol = opt->xatol;
  p.fatol = opt->fatol;
  // SciPy's defaults (_minimize_neldermead): both limits missing -> N * 200 each; one missing ->
  // unlimited, unless the other one is unlimited too
  long long maxiter = opt->maxiter, maxfev = opt->maxfev;  // < 0: not given; LLONG_MAX: infinite
  if (maxiter < 0 && maxfev < 0) maxiter = maxfev = (long long)nv * 200;
  else if (maxiter < 0) maxiter = (maxfev == LLONG_MAX) ? (long long)nv * 200 : LLONG_MAX;
  else if (maxfev < 0) maxfev = (maxiter == LLONG_MAX) ? (long long)nv * 200 : LLONG_MAX;
  p.maxiter = maxiter;
  p.maxfev = maxfev;
  p.adaptive = opt->adaptive != 0;
  p.out = reinterpret_cast<double*>(w + o_out);
  p.out_stride = out_stride;
  p.gws = global_staging ? reinterpret_cast<float*>(w + o_gws) : nullptr;

  const unsigned grid = (unsigned)std::min<int64_t>(n_patterns, global_staging ? max_ctas : 0x7fffffff);
  auto launch = [&](auto kernel) -> int {
    KDI_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    kernel<<<grid, kRefThreads, smem, st>>>(p);
    KDI_CUDA(ctx, cudaGetLastError());
    KDI_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    return KDI_OK;
  };
  // resident CTAs per SM the register allocation aims for (KDI_REFINE_MINB = 2, 3 or 4: tuning aid)
  static const int minb = [] { const char* e = getenv("KDI_REFINE_MINB"); return e ? atoi(e) : 4; }();
#define KDI_REFINE_LAUNCH(MB)                                                   \
  do {                                                                          \
    if (mode == KDI_REFINE_ORI) KDI_TRY(launch(kdi_refine_kernel<0, 3, MB>));   \
    else if (mode == KDI_REFINE_PC) KDI_TRY(launch(kdi_refine_kernel<1, 3, MB>)); \
    else KDI_TRY(launch(kdi_refine_kernel<2, 6, MB>));                          \
  } while (0)
  if (objective) {
    if (mode == KDI_REFINE_ORI) KDI_TRY(launch(kdi_refine_objective_kernel<0, 3>));
    else if (mode == KDI_REFINE_PC) KDI_TRY(launch(kdi_refine_objective_kernel<1, 3>));
    else KDI_TRY(launch(kdi_refine_objective_kernel<2, 6>));
  } else if (minb <= 2) KDI_REFINE_LAUNCH(2);
  else if (minb == 3) KDI_REFINE_LAUNCH(3);
  else KDI_REFINE_LAUNCH(4);
#undef KDI_REFINE_LAUNCH
  ctx->tm.kernel_launches++;
  KDI_CUDA(ctx, cudaMemcpyAsync(results_out, w + o_out, (size_t)n_patterns * out_stride * 8, cudaMemcpyDeviceToHost, st));
  ctx->tm.d2h_bytes += n_patterns * out_stride * 8;
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  KDI_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.total_ms = ms;
  return KDI_OK;
}

extern "C" int kdi_refine(kdi_ctx* ctx, const kdi_master_pattern* mp, int mode, const void* patterns,
                          int pat_loc, int pat_dtype, int64_t n_patterns, int nrows, int ncols, int rescale,
                          const double* x0, int n_starts, const double* lower, const double* upper,
                          const double* rotations, const double* pcs, const double* om_detector_to_sample,
                          const kdi_refine_options* opt, double* results_out) {
  return refine_impl(ctx, mp, mode, patterns, pat_loc, pat_dtype, n_patterns, nrows, ncols, rescale, x0, n_starts,
                     lower, upper, rotations, pcs, om_detector_to_sample, opt, results_out, false, nullptr, 0);
}

extern "C" int kdi_refine_objective(kdi_ctx* ctx, const kdi_master_pattern* mp, int mode, const void* patterns,
                                    int pat_loc, int pat_dtype, int64_t n_source_patterns, int nrows, int ncols,
                                    int rescale, const int64_t* pattern_rows, int64_t n_rows, const double* x,
                                    int n_points, const double* rotations, const double* pcs,
                                    const double* om_detector_to_sample, double* values_out) {
  return refine_impl(ctx, mp, mode, patterns, pat_loc, pat_dtype, n_rows, nrows, ncols, rescale, x, n_points, nullptr,
                     nullptr, rotations, pcs, om_detector_to_sample, nullptr, values_out, true, pattern_rows,
                     n_source_patterns);
}
