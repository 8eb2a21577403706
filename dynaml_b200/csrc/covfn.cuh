// covfn.cuh -- device-side covariance functions and their hyper-parameter derivatives.
// Each formula restates the reference's scalar closure (file:line in the comments); the
// host-side make_cov_params() folds the hyper-parameters into the constants used here.
#pragma once
#include <math.h>
#include "common.cuh"

namespace dgp {

constexpr int MAX_MATERN_P = 8;

struct CompositeDev;

struct CovParams {
  const CompositeDev *comp = nullptr;  // host pointer, DGP_KERNEL_COMPOSITE only (copied by value at launch)
  int kind = 0;
  int p = 0;           // Matern order
  int grad_mode = 0;
  int l1 = 0;          // 1: the kernel consumes |x-y|_1 (periodic), else |x-y|_2^2
  double a2 = 1.0;     // SE/ARD: amplitude^2, periodic: sigma^2, Dirac: |s|, else 1
  double c1 = 0.0;     // SE/RBF: 1/(2 l^2); ARD: 0.5 (inputs pre-scaled by 1/b_i); Matern: sqrt(2nu)/l;
                       // periodic: pi*frequency
  double c2 = 0.0;     // Matern: sqrt(8nu)/l; periodic: 2/l^2
  double c3 = 0.0;     // Matern: l; periodic: pi*frequency/l^2 (the reference's gradient "k" factor)
  double coef[MAX_MATERN_P + 1];  // Matern: Gamma(p+1)/Gamma(2p+1) * (p+i)!/(i!(p-i)!), i = 0..p
  double noise_abs = 0.0;         // |noiseLevel| of the Dirac noise model, added where |x-y|_2 == 0
};

// Value of the covariance from the pair distance (`dist` = |x-y|^2, or |x-y|_1 when cp.l1).
template <int KIND>
__device__ __forceinline__ double cov_value(const CovParams &cp, double dist) {
  if (KIND == DGP_KERNEL_SE || KIND == DGP_KERNEL_RBF || KIND == DGP_KERNEL_ARD_SE) {
    // RBFKernel.scala:42-43, 79-80, 131-132:  a^2 exp(-|x-y|^2 / (2 l^2))
    return cp.a2 * exp(-dist * cp.c1);
  } else if (KIND == DGP_KERNEL_MATERN) {
    // GenericMaternKernel.scala:55-68
    double r = sqrt(dist);
    double u = cp.c2 * r;
    double s = cp.coef[0];
    for (int i = 1; i <= cp.p; ++i) s = s * u + cp.coef[i];
    return exp(-cp.c1 * r) * s;
  } else if (KIND == DGP_KERNEL_PERIODIC) {
    // PeriodicKernel.scala:34-46:  sigma^2 exp(-2 sin^2(pi f |x-y|_1) / l^2)
    double sn = sin(dist * cp.c1);
    return cp.a2 * exp(-cp.c2 * sn * sn);
  } else if (KIND == DGP_KERNEL_CAUCHY) {
    // CauchyKernel.scala: 1/(1 + (|x-y|/sigma)^2)
    return 1.0 / fma(dist, cp.c1, 1.0);
  } else if (KIND == DGP_KERNEL_LAPLACIAN) {
    // LaplacianKernel.scala: exp(-|x-y|_1 / beta)
    return exp(-dist * cp.c1);
  } else if (KIND == DGP_KERNEL_RATQUAD) {
    // RationalQuadraticKernel.scala: (1 + |x-y|^2/(mu l^2))^(-(d + mu)/2)
    return pow(fma(dist, cp.c1, 1.0), cp.c2);
  } else if (KIND == DGP_KERNEL_TSTUDENT) {
    // TStudentKernel.scala: 1/(1 + |x-y|^d)
    return 1.0 / (1.0 + pow(sqrt(dist), cp.c1));
  } else {
    // DiracKernel.scala:43-47
    return dist == 0.0 ? cp.a2 : 0.0;
  }
}

// Kinds whose distance is |x-y|_1 instead of |x-y|_2^2.
__host__ __device__ constexpr bool kind_is_l1(int kind) {
  return kind == DGP_KERNEL_PERIODIC || kind == DGP_KERNEL_LAPLACIAN;
}

// Kinds that are not functions of x - y: they consume x.y, |x|^2, |y|^2 of the original points.
__host__ __device__ constexpr bool kind_is_inner_product(int kind) {
  return kind == DGP_KERNEL_POLYNOMIAL || kind == DGP_KERNEL_FBM;
}

// Pair geometry a composite evaluation has at hand.
struct PairGeom {
  double r2 = 0.0, l1 = 0.0, dot = 0.0, nx = 0.0, ny = 0.0;
};

__device__ __forceinline__ double cov_value_ip(const CovParams &cp, const PairGeom &q) {
  if (cp.kind == DGP_KERNEL_POLYNOMIAL) {
    // PolynomialKernel.scala:30-33: math.pow((x.t * y) + offset, degree.toInt)
    return pow(q.dot + cp.c1, (double)cp.p);
  }
  // FBMKernel.scala:84-91: 0.5 (|x|^2H + |y|^2H - |x-y|^2H) with the norms themselves raised to 2H
  return 0.5 * (pow(sqrt(q.nx), cp.c2) + pow(sqrt(q.ny), cp.c2) - pow(sqrt(q.r2), cp.c2));
}

// Runtime-dispatched value (composite kernels evaluate several kinds per pair; the switch is warp-uniform).
__device__ __forceinline__ double cov_value_rt(const CovParams &cp, double dist) {
  switch (cp.kind) {
    case DGP_KERNEL_SE:
    case DGP_KERNEL_RBF:
    case DGP_KERNEL_ARD_SE: return cov_value<DGP_KERNEL_SE>(cp, dist);
    case DGP_KERNEL_MATERN: return cov_value<DGP_KERNEL_MATERN>(cp, dist);
    case DGP_KERNEL_PERIODIC: return cov_value<DGP_KERNEL_PERIODIC>(cp, dist);
    case DGP_KERNEL_CAUCHY: return cov_value<DGP_KERNEL_CAUCHY>(cp, dist);
    case DGP_KERNEL_LAPLACIAN: return cov_value<DGP_KERNEL_LAPLACIAN>(cp, dist);
    case DGP_KERNEL_RATQUAD: return cov_value<DGP_KERNEL_RATQUAD>(cp, dist);
    case DGP_KERNEL_TSTUDENT: return cov_value<DGP_KERNEL_TSTUDENT>(cp, dist);
    default: return cov_value<DGP_KERNEL_DIRAC>(cp, dist);
  }
}

// One out-of-line copy of the kind switch for the composite kernels: their factor loops are unrolled, and four
// inlined copies of every transcendental path made the kernels instruction-cache bound (ncu: `no_instruction`
// was the second largest stall).  The covariance constants travel by value; Matern orders above 3 (coefficients
// do not fit the four scalars) and the inner-product kinds take the inline paths.
static __device__ __noinline__ double cov_value_call(int kind, double dist, double a2, double c1, double c2, int p, double k0,
                                              double k1, double k2, double k3) {
  switch (kind) {
    case DGP_KERNEL_SE:
    case DGP_KERNEL_RBF:
    case DGP_KERNEL_ARD_SE: return a2 * exp(-dist * c1);
    case DGP_KERNEL_MATERN: {
      const double r = sqrt(dist), u = c2 * r;
      double s = k0;
      if (p >= 1) s = s * u + k1;
      if (p >= 2) s = s * u + k2;
      if (p >= 3) s = s * u + k3;
      return exp(-c1 * r) * s;
    }
    case DGP_KERNEL_PERIODIC: {
      const double sn = sin(dist * c1);
      return a2 * exp(-c2 * sn * sn);
    }
    case DGP_KERNEL_CAUCHY: return 1.0 / fma(dist, c1, 1.0);
    case DGP_KERNEL_LAPLACIAN: return exp(-dist * c1);
    case DGP_KERNEL_RATQUAD: return pow(fma(dist, c1, 1.0), c2);
    case DGP_KERNEL_TSTUDENT: return 1.0 / (1.0 + pow(sqrt(dist), c1));
    default: return dist == 0.0 ? a2 : 0.0;
  }
}

__device__ __forceinline__ double cov_value_shared(const CovParams &cp, double dist) {
  if (cp.kind == DGP_KERNEL_MATERN && cp.p > 3) return cov_value<DGP_KERNEL_MATERN>(cp, dist);
  return cov_value_call(cp.kind, dist, cp.a2, cp.c1, cp.c2, cp.p, cp.coef[0], cp.coef[1], cp.coef[2], cp.coef[3]);
}

// ---- kernel algebra: k = sum_t coef_t prod_f v_f^e_tf (DGP_KERNEL_COMPOSITE) -------------------------
constexpr int MAXF = DGP_MAX_FACTORS, MAXT = DGP_MAX_TERMS, MAX_ARD_SLOTS = 4, MAX_ARD_DIM = 32;

// Everything the device needs, passed to the kernels by value (~1.7 KB of parameter space).
struct CompositeDev {
  int F = 0, T = 0;
  int any_l1 = 0;                 // some factor consumes |x-y|_1
  int any_ip = 0;                 // some factor consumes x.y / |x|^2 / |y|^2 (the points are then the raw ones)
  int nslots = 0;                 // weighted-metric slots in use
  int simple = 0;                 // every exponent is 0 or 1: terms are products over the bit sets mask[t]
  unsigned mask[MAXT];            // bit f set: base kernel f is a factor of term t
  int metric[MAXF];               // per factor: 0 = |x-y|_2^2, 1 = |x-y|_1, 2 + s = weighted slot s
  int slot_l1[MAX_ARD_SLOTS];     // slot s accumulates sum_c w_c |x_c - y_c| instead of sum_c w_c (x_c - y_c)^2
  int dim0[MAXF], dimn[MAXF];     // coordinates [dim0, dim0 + dimn) a factor sees (tensor kernels); whole vector by default
  int moment_off[MAXF + 1];       // gradient pass: first moment of each factor; [F] = the noise moment
  int ard_exact[MAXF];            // ARD factor with the exact per-dimension derivative
  CovParams cp[MAXF];
  double coef[MAXT];
  unsigned char expo[MAXT][MAXF];
  double w[MAX_ARD_SLOTS][MAX_ARD_DIM];  // slot weights: 1 / b_c^2 (ARD-SE) and / or a 0-1 coordinate mask
};

__device__ __forceinline__ double ipow_small(double v, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= v;
  return r;
}

// The composite kernels are instantiated for FT = 4 (expressions of up to four base kernels: every loop over the
// factors is unrolled and v / cof live in registers) and FT = MAXF (runtime loops, local-memory arrays).
template <int FT>
struct FactorLoop {
  static constexpr bool unrolled = FT <= 4;
};

// Value of the expression from the base-kernel values v[0..F).
template <int FT>
__device__ __forceinline__ double composite_combine(const CompositeDev &c, const double (&v)[FT]) {
  double k = 0.0;
  if (FactorLoop<FT>::unrolled && c.simple) {  // branch-free: the usual sums / products of distinct kernels
    for (int t = 0; t < c.T; ++t) {
      const unsigned mk = c.mask[t];
      double prod = c.coef[t];
#pragma unroll
      for (int f = 0; f < FT; ++f) prod *= ((mk >> f) & 1u) ? v[f] : 1.0;
      k += prod;
    }
    return k;
  }
  for (int t = 0; t < c.T; ++t) {
    double prod = c.coef[t];
    if (FactorLoop<FT>::unrolled) {
#pragma unroll
      for (int f = 0; f < FT; ++f) {
        const int e = (f < c.F) ? c.expo[t][f] : 0;
        if (e) prod *= ipow_small(v[f], e);
      }
    } else {
      for (int f = 0; f < c.F; ++f) {
        const int e = c.expo[t][f];
        if (e) prod *= ipow_small(v[f], e);
      }
    }
    k += prod;
  }
  return k;
}

// cof[f] = d k / d v_f for every base kernel: the factor each derivative of base kernel f is multiplied by
// (MultiplicativeCovariance.scala:68-77; `* c` scales value and gradient alike, LocalScalarKernel.scala:81-85).
// One pass per term with running prefix / suffix products: O(T F) per pair and no division by v_f.
template <int FT>
__device__ __forceinline__ void composite_cofactors(const CompositeDev &c, const double (&v)[FT], double (&cof)[FT]) {
  if (FactorLoop<FT>::unrolled && c.simple) {
#pragma unroll
    for (int f = 0; f < FT; ++f) cof[f] = 0.0;
    for (int t = 0; t < c.T; ++t) {
      const unsigned mk = c.mask[t];
      double u[FT], suffix[FT + 1];
#pragma unroll
      for (int f = 0; f < FT; ++f) u[f] = ((mk >> f) & 1u) ? v[f] : 1.0;
      suffix[FT] = 1.0;
#pragma unroll
      for (int f = FT - 1; f >= 0; --f) suffix[f] = suffix[f + 1] * u[f];
      double prefix = c.coef[t];
#pragma unroll
      for (int f = 0; f < FT; ++f) {
        cof[f] += ((mk >> f) & 1u) ? prefix * suffix[f + 1] : 0.0;
        prefix *= u[f];
      }
    }
    return;
  }
  if (FactorLoop<FT>::unrolled) {
#pragma unroll
    for (int f = 0; f < FT; ++f) cof[f] = 0.0;
    for (int t = 0; t < c.T; ++t) {
      double suffix = 1.0, tmp[FT];
#pragma unroll
      for (int f = FT - 1; f >= 0; --f) {
        tmp[f] = suffix;
        const int e = (f < c.F) ? c.expo[t][f] : 0;
        if (e) suffix *= ipow_small(v[f], e);
      }
      double prefix = c.coef[t];
#pragma unroll
      for (int f = 0; f < FT; ++f) {
        const int e = (f < c.F) ? c.expo[t][f] : 0;
        if (e) {
          cof[f] = fma((double)e * ipow_small(v[f], e - 1), prefix * tmp[f], cof[f]);
          prefix *= ipow_small(v[f], e);
        }
      }
    }
  } else {
    for (int f = 0; f < c.F; ++f) cof[f] = 0.0;
    for (int t = 0; t < c.T; ++t) {
      double suffix = 1.0, tmp[FT];
      for (int f = c.F - 1; f >= 0; --f) {
        tmp[f] = suffix;  // product over g > f
        const int e = c.expo[t][f];
        if (e) suffix *= ipow_small(v[f], e);
      }
      double prefix = c.coef[t];
      for (int f = 0; f < c.F; ++f) {
        const int e = c.expo[t][f];
        if (e) {
          cof[f] = fma((double)e * ipow_small(v[f], e - 1), prefix * tmp[f], cof[f]);
          prefix *= ipow_small(v[f], e);
        }
      }
    }
  }
}

// Host: fold a Spec into CovParams.  `noise` < 0 is legal (the reference takes abs()).
inline CovParams make_cov_params(const Spec &s, bool with_noise, int d = 0);

// Parsed composite: the device block plus the base-kernel specs (for grad_finish).
struct Composite {
  CompositeDev dev;
  Spec factor[MAXF];
  int factor_d[MAXF];  // feature count each base kernel sees (its slice of the input vector)
};

inline CovParams make_cov_params(const Spec &s, bool with_noise, int d) {
  CovParams cp;
  cp.kind = s.kind;
  cp.grad_mode = s.grad_mode;
  cp.noise_abs = with_noise ? fabs(s.noise) : 0.0;
  for (int i = 0; i <= MAX_MATERN_P; ++i) cp.coef[i] = 0.0;
  switch (s.kind) {
    case DGP_KERNEL_SE:
      cp.a2 = s.hyp[1] * s.hyp[1];
      cp.c1 = 1.0 / (2.0 * s.hyp[0] * s.hyp[0]);
      break;
    case DGP_KERNEL_RBF:
      cp.a2 = 1.0;
      cp.c1 = 1.0 / (2.0 * s.hyp[0] * s.hyp[0]);
      break;
    case DGP_KERNEL_ARD_SE:
      cp.a2 = s.hyp[0] * s.hyp[0];
      cp.c1 = 0.5;
      break;
    case DGP_KERNEL_MATERN: {
      int p = (int)floor(fabs(s.hyp[0]));  // setHyperParameters floors |p| (GenericMaternKernel.scala:48-53)
      double nu = p + 0.5, l = s.hyp[1];
      cp.p = p;
      cp.c1 = sqrt(2.0 * nu) / l;
      cp.c2 = sqrt(8.0 * nu) / l;
      cp.c3 = l;
      double lead = exp(lgamma((double)p + 1.0) - lgamma(2.0 * p + 1.0));
      auto fact = [](int n) { double f = 1.0; for (int i = 2; i <= n; ++i) f *= i; return f; };
      for (int i = 0; i <= p; ++i) cp.coef[i] = lead * (fact(p + i) / (fact(i) * fact(p - i)));
      break;
    }
    case DGP_KERNEL_PERIODIC:
      cp.l1 = 1;
      cp.a2 = s.hyp[0] * s.hyp[0];
      cp.c1 = M_PI * s.hyp[2];
      cp.c2 = 2.0 / (s.hyp[1] * s.hyp[1]);
      cp.c3 = M_PI * s.hyp[2] / (s.hyp[1] * s.hyp[1]);
      break;
    case DGP_KERNEL_DIRAC:
      cp.a2 = fabs(s.hyp[0]);
      break;
    case DGP_KERNEL_CAUCHY:
      cp.c1 = 1.0 / (s.hyp[0] * s.hyp[0]);
      break;
    case DGP_KERNEL_LAPLACIAN:
      cp.l1 = 1;
      cp.c1 = 1.0 / s.hyp[0];
      break;
    case DGP_KERNEL_RATQUAD:  // hyp = [mu, l]; the exponent uses the feature count x.length
      cp.c1 = 1.0 / (s.hyp[0] * s.hyp[1] * s.hyp[1]);
      cp.c2 = -0.5 * ((double)d + s.hyp[0]);
      cp.c3 = 1.0 / ((s.hyp[0] * s.hyp[1]) * (s.hyp[0] * s.hyp[1]));
      break;
    case DGP_KERNEL_TSTUDENT:
      cp.c1 = s.hyp[0];
      break;
    case DGP_KERNEL_COMPOSITE:
      cp.comp = s.comp ? &s.comp->dev : nullptr;
      break;
    case DGP_KERNEL_POLYNOMIAL:
      cp.p = (int)s.hyp[0];  // config("degree").toInt
      cp.c1 = s.hyp[1];
      break;
    case DGP_KERNEL_FBM:
      cp.c1 = s.hyp[0];
      cp.c2 = 2.0 * s.hyp[0];
      break;
  }
  return cp;
}

}  // namespace dgp
