// grad.cu -- the hyper-parameter gradient traces of gradEnergy without materialising dK/dtheta.
//
// Reference (CORE/models/gp/AbstractGPRegressionModel.scala:231-241): for every theta it builds
// the full N x N matrix dK/dtheta (SVMKernel.scala:251-307), multiplies alpha alpha^T by it (an
// N^3 GEMM), runs two N x N multi-RHS block solves, and takes the trace.  Algebraically
//     g_theta = sum_ij (alpha_i alpha_j - (K^-1)_ij) * dK_ij/dtheta ,
// so one streaming pass over the lower tiles of K^-1 that re-derives dK_ij/dtheta from the pair
// distance in registers gives every g_theta at once.  HBM-read bound (8 bytes per pair).
//
// The kernel accumulates kind-specific "moments" S_q = sum_ij M_ij * f_q(x_i, x_j); grad_finish()
// maps them to the reference's per-hyper-parameter derivatives on the host.
#include "grad.cuh"
#include "lean.cuh"

namespace dgp {

int grad_partials_per_tile(const CovParams &cp, int dp);

namespace {

constexpr int GT = 128;
constexpr int GTHREADS = 256;

enum Mode { M_SE = 0, M_ARD_REF = 1, M_ARD_EXACT = 2, M_MATERN = 3, M_PERIODIC = 4, M_DIRAC = 5, M_CAUCHY = 6,
            M_LAPLACIAN = 7, M_RATQUAD = 8, M_TSTUDENT = 9, M_POLY = 10, M_FBM = 11 };

template <int MODE, int DP>
struct NSums {
  // [last] is always the noise moment sum M_ij 1[|x_i - x_j| == 0]
  static constexpr int value = (MODE == M_SE) ? 3 : (MODE == M_ARD_REF) ? 3 : (MODE == M_ARD_EXACT) ? (DP + 2)
                               : (MODE == M_MATERN) ? 2 : (MODE == M_PERIODIC) ? 3 : (MODE == M_RATQUAD) ? 3 : 2;
};

// Per-pair contribution.  `dist` is |x-y|^2 (for ARD: of the pre-scaled inputs), `plain2` the
// unscaled |x-y|^2 (ARD reference mode), `sq[c]` the unscaled squared differences (ARD exact).
template <int MODE, int DP>
__device__ __forceinline__ void accumulate(const CovParams &cp, double m, double dist, double plain2, const double *sq,
                                           double *s) {
  constexpr int NS = NSums<MODE, DP>::value;
  if (dist == 0.0) s[NS - 1] += m;
  if (MODE == M_SE) {
    // RBFKernel.scala:45-52, 82-87: d/dl = k |x-y|^2 / |l|^3 ; d/da = 2 a exp(.)
    double e = exp(-dist * cp.c1);
    s[0] = fma(m * e, dist, s[0]);
    s[1] = fma(m, e, s[1]);
  } else if (MODE == M_ARD_REF) {
    // RBFKernel.scala:134-145: d/dh = 2 k / h ; d/db_i = k |x-y|_2^2 / b_i^3 (same for every i, defect D2)
    double e = exp(-dist * cp.c1);
    s[0] = fma(m, e, s[0]);
    s[1] = fma(m * e, plain2, s[1]);
  } else if (MODE == M_ARD_EXACT) {
    double e = exp(-dist * cp.c1);
    s[0] = fma(m, e, s[0]);
    double me = m * e;
#pragma unroll
    for (int c = 0; c < DP; ++c) s[1 + c] = fma(me, sq[c], s[1 + c]);
  } else if (MODE == M_MATERN) {
    // GenericMaternKernel.scala:70-97: dk/dl = diffLead*sumTerm + diffSum*leadingTerm
    double r = sqrt(dist);
    double u = cp.c2 * r;
    double sum = cp.coef[0], q = (double)cp.p * cp.coef[0];
    for (int i = 1; i <= cp.p; ++i) {
      sum = sum * u + cp.coef[i];
      q = q * u + (double)(cp.p - i) * cp.coef[i];
    }
    double lead = exp(-cp.c1 * r);
    // q currently holds sum_i (p-i) coef_i u^(p-i) / u^0 evaluated by Horner on the same powers
    double dk = (cp.c1 * r * (lead * sum) - lead * q) / cp.c3;
    s[0] = fma(m, dk, s[0]);
  } else if (MODE == M_PERIODIC) {
    // PeriodicKernel.scala:48-62
    double sn = sin(dist * cp.c1);
    double k = cp.a2 * exp(-cp.c2 * sn * sn);
    double kap = cp.c3 * dist;
    s[0] = fma(m, k, s[0]);
    s[1] = fma(m * k, sin(2.0 * kap) * kap, s[1]);
  } else if (MODE == M_CAUCHY) {
    // CauchyKernel.scala: d/dsigma = 2 k^2 |x-y|^2 / sigma^3
    double k = 1.0 / fma(dist, cp.c1, 1.0);
    s[0] = fma(m * k * k, dist, s[0]);
  } else if (MODE == M_LAPLACIAN) {
    // LaplacianKernel.scala: d/dbeta = k |x-y|_1 / beta^2
    double k = exp(-dist * cp.c1);
    s[0] = fma(m * k, dist, s[0]);
  } else if (MODE == M_RATQUAD) {
    // RationalQuadraticKernel.scala: grad_l = -2 e r^2 base^(e-1) / (mu l^3);
    //                                grad_mu = k * ( -e r^2 / (base (mu l)^2) - 0.5 log(base) )
    double base = fma(dist, cp.c1, 1.0);
    double k = pow(base, cp.c2);
    s[0] = fma(m * dist, k / base, s[0]);
    s[1] = fma(m * k, -cp.c2 * dist * cp.c3 / base - 0.5 * log(base), s[1]);
  } else if (MODE == M_TSTUDENT) {
    // TStudentKernel.scala: d/dd = -r^d log(r) / (1 + r^d)^2 for r > 0, else 0
    if (dist > 0.0) {
      double r = sqrt(dist);
      double rd = pow(r, cp.c1);
      s[0] = fma(m, -rd * log(r) / ((1.0 + rd) * (1.0 + rd)), s[0]);
    }
  } else {
    // DiracKernel.scala:49-53: value / |s|
    if (dist == 0.0) s[0] += m;
  }
}

template <int MODE, int DP>
__global__ void __launch_bounds__(GTHREADS) grad_trace_kernel(CovParams cp, GradArgs a) {
  constexpr int NS = NSums<MODE, DP>::value;
  extern __shared__ __align__(16) double sm[];
  // layout: column points (GT x DP) | scaled column points (ARD only) | alpha_j (GT) | reduction (8 x NS)
  constexpr bool ARD = (MODE == M_ARD_REF || MODE == M_ARD_EXACT);
  double *sx = sm;
  double *sxs = sm + GT * DP;
  double *sal = sm + (ARD ? 2 : 1) * GT * DP;
  double *red = sal + GT;

  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;
  int64_t tlin = blockIdx.x;
  int ti, tj;
  if (a.edge_only) {
    // the tiles the lean pass leaves out: the mt diagonal tiles, then the rest of the last tile row (padding)
    const int mt = (int)(a.np / GT);
    ti = tlin < mt ? (int)tlin : mt - 1;
    tj = tlin < mt ? (int)tlin : (int)(tlin - mt);
    tlin = (int64_t)ti * (ti + 1) / 2 + tj;
  } else {
    ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
    while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
    tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  }
  const int64_t m0 = (int64_t)ti * GT, n0 = (int64_t)tj * GT;

  {
    const double2 *src = reinterpret_cast<const double2 *>(a.X + n0 * DP);
    double2 *dst = reinterpret_cast<double2 *>(sx);
    for (int i = tid; i < GT * DP / 2; i += GTHREADS) dst[i] = src[i];
    if (ARD) {
      const double2 *src2 = reinterpret_cast<const double2 *>(a.Xs + n0 * DP);
      double2 *dst2 = reinterpret_cast<double2 *>(sxs);
      for (int i = tid; i < GT * DP / 2; i += GTHREADS) dst2[i] = src2[i];
    }
    if (tid < GT) sal[tid] = a.alpha[n0 + tid];
  }

  const int64_t r0 = m0 + 2 * tx;
  double x0[DP], x1[DP];
  double xs0[ARD ? DP : 1], xs1[ARD ? DP : 1];
#pragma unroll
  for (int c = 0; c < DP; ++c) {
    x0[c] = a.X[r0 * DP + c];
    x1[c] = a.X[(r0 + 1) * DP + c];
  }
  if (ARD) {
#pragma unroll
    for (int c = 0; c < DP; ++c) {
      xs0[c] = a.Xs[r0 * DP + c];
      xs1[c] = a.Xs[(r0 + 1) * DP + c];
    }
  }
  const double al0 = a.alpha[r0], al1 = a.alpha[r0 + 1];
  __syncthreads();

  double s[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) s[q] = 0.0;

  const double *kinv = a.Kinv + r0 + n0 * a.ld;
  for (int cc = 0; cc < GT / 4; ++cc) {
    const int j = ty + 4 * cc;
    const int64_t col = n0 + j;
    const double2 kv = *reinterpret_cast<const double2 *>(kinv + (int64_t)j * a.ld);
    const double aj = sal[j];
    double d0 = 0.0, d1 = 0.0, p0 = 0.0, p1 = 0.0;
    double sq0[MODE == M_ARD_EXACT ? DP : 1], sq1[MODE == M_ARD_EXACT ? DP : 1];
#pragma unroll
    for (int c = 0; c < DP; ++c) {
      const double y = sx[j * DP + c];
      const double e0 = x0[c] - y, e1 = x1[c] - y;
      if (MODE == M_PERIODIC || MODE == M_LAPLACIAN) {
        d0 += fabs(e0);
        d1 += fabs(e1);
      } else if (ARD) {
        const double ys = sxs[j * DP + c];
        const double f0 = xs0[c] - ys, f1 = xs1[c] - ys;
        d0 = fma(f0, f0, d0);
        d1 = fma(f1, f1, d1);
        if (MODE == M_ARD_EXACT) {
          sq0[c] = e0 * e0;
          sq1[c] = e1 * e1;
        } else {
          p0 = fma(e0, e0, p0);
          p1 = fma(e1, e1, p1);
        }
      } else {
        d0 = fma(e0, e0, d0);
        d1 = fma(e1, e1, d1);
      }
    }
    // weights: strictly-lower pairs count twice (symmetry), the diagonal once, everything
    // above the diagonal or in the padding not at all
    double w0 = (r0 > col) ? 2.0 : (r0 == col ? 1.0 : 0.0);
    double w1 = (r0 + 1 > col) ? 2.0 : (r0 + 1 == col ? 1.0 : 0.0);
    if (r0 >= a.n || col >= a.n) w0 = 0.0;
    if (r0 + 1 >= a.n || col >= a.n) w1 = 0.0;
    if (w0 != 0.0) accumulate<MODE, DP>(cp, w0 * (al0 * aj - kv.x), d0, p0, sq0, s);
    if (w1 != 0.0) accumulate<MODE, DP>(cp, w1 * (al1 * aj - kv.y), d1, p1, sq1, s);
  }

  // ---- deterministic block reduction --------------------------------------------------------
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = s[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NS + q] = v;
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
    for (int wdx = 0; wdx < GTHREADS / 32; ++wdx) v += red[wdx * NS + tid];
    a.partials[tlin * NS + tid] = v;
  }
}

// sums[q] = sum over tiles of partials[tile][q], fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *partials, int64_t tiles, int ns,
                                                               double *sums) {
  __shared__ double sh[1024];
  const int q = blockIdx.x;
  double v = 0.0;
  for (int64_t t = threadIdx.x; t < tiles; t += blockDim.x) v += partials[t * ns + q];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[q] = sh[0];
}

// ---------------------------------------------------------------------------------------------------------------
// Lean trace pass: the strictly-lower tiles of a training set whose points are pairwise distinct (no pair at
// distance zero, so no noise moment there) under the accuracy gate of the Gram assembly (gram_is_accurate).  Same
// structure as assemble_gram_lean_kernel: the scaled squared distance of 32 x 16 pairs per warp and pass comes out of
// m8n8k4 DMMAs (|x|^2 + |y|^2 - 2 x.y), exp / sqrt are the short table / rsqrt-seed sequences of lean.cuh, and the
// entry of K^-1 arrives as a 16-byte load issued before the DMMAs.  Every pair of these tiles has weight 2
// (symmetry), applied once to the tile's partial sums.  Two CTAs per SM: the 16 prefetched K^-1 entries and the 16
// accumulators per thread do not fit the 80 registers of three (measured on B200, profiles/trace_lean_ab_r2.txt:
// 2.27 ms with the spills against 1.83 ms, Matern-5/2 d = 16 N = 32768; the exact-difference pass takes 6.01 ms).
//   SE       S0 += m e acc (acc = c1 r^2; the tile sum is divided by c1), S1 += m e
//   ARD-SE   S0 += m e, S1 += m e |x - y|^2 of the unscaled points (reference mode, D2): a second Gram product
//   Matern   dk/dl = exp(-rho) rho Q(rho), rho = sqrt(2 nu) r / l: the reference's diffLead * sumTerm + diffSum *
//            leadingTerm (GenericMaternKernel.scala:70-97) collected by powers of rho on the host; its constant term
//            is identically zero, Q has degree p
// The diagonal tiles and the tiles touching the padding go through grad_trace_kernel (edge_only).
struct LeanTraceCov {
  double scale;            // operands are multiplied by this when staged
  double q0, q1, q2, q3;   // Matern: Q(rho), right-aligned (q3 is the constant term)
  double s0_scale;         // SE: 1 / c1
};

template <int MODE, int DP, bool P3>
__global__ void __launch_bounds__(GTHREADS, 2)
    grad_trace_lean_kernel(LeanTraceCov f, LeanConst lc, ExpTab T, GradArgs a) {
  constexpr int NS = NSums<MODE, DP>::value;
  constexpr bool ARD = (MODE == M_ARD_REF);
  constexpr int LDK = lean_ldk<DP>();
  extern __shared__ __align__(16) double sm_raw[];
  double *tab = sm_raw;                    // 64 entries x 16 copies
  double *Brow = tab + 64 * 16;            // row points i (MMA N): [i][k] = scale * x
  double *Acol = Brow + GT * LDK;          // column points j (MMA M): [j][k] = -2 scale * y
  double *nx = Acol + GT * LDK;            // scale^2 |x_i|^2
  double *ny = nx + GT;
  double *sal = ny + GT;                   // alpha of the rows, then of the columns
  double *red = sal + 2 * GT;              // 8 x NS
  double *Brow2 = red + 8 * NS;            // ARD: the unscaled points, same layout (8 * NS is even: 16-byte aligned)
  double *Acol2 = Brow2 + GT * LDK;
  double *nx2 = Acol2 + GT * LDK;
  double *ny2 = nx2 + GT;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;

  // strictly-lower tile (ti, tj), tj < ti: the triangular numbering of (ti - 1, tj)
  const int64_t tl = blockIdx.x;
  int tr = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
  while ((int64_t)tr * (tr + 1) / 2 > tl) --tr;
  while ((int64_t)(tr + 1) * (tr + 2) / 2 <= tl) ++tr;
  const int tj = (int)(tl - (int64_t)tr * (tr + 1) / 2), ti = tr + 1;
  const int64_t m0 = (int64_t)ti * GT, n0 = (int64_t)tj * GT;
  if (m0 + GT > a.n) return;   // touches the padding: edge pass
  const int64_t tlin = (int64_t)ti * (ti + 1) / 2 + tj;

  {
    const bool is_row = tid < GT;
    const int pidx = is_row ? tid : tid - GT;
    const int64_t pt = (is_row ? m0 : n0) + pidx;
    const double *src = (ARD ? a.Xs : a.X) + pt * DP;
    double *dst = (is_row ? Brow : Acol) + pidx * LDK;
    const double sc = is_row ? 1.0 : -2.0;
    double v[DP];
#pragma unroll
    for (int c = 0; c < DP / 2; ++c) {
      const double2 q = reinterpret_cast<const double2 *>(src)[c];
      v[2 * c] = q.x;
      v[2 * c + 1] = q.y;
    }
    double nrm = 0.0;
#pragma unroll
    for (int c = 0; c < DP; ++c) {
      v[c] *= f.scale;
      nrm = fma(v[c], v[c], nrm);
    }
#pragma unroll
    for (int c = 0; c < DP / 2; ++c)
      reinterpret_cast<double2 *>(dst)[c] = make_double2(sc * v[2 * c], sc * v[2 * c + 1]);
    (is_row ? nx : ny)[pidx] = nrm;
    if (ARD) {
      const double *src2 = a.X + pt * DP;
      double *dst2 = (is_row ? Brow2 : Acol2) + pidx * LDK;
#pragma unroll
      for (int c = 0; c < DP / 2; ++c) {
        const double2 q = reinterpret_cast<const double2 *>(src2)[c];
        v[2 * c] = q.x;
        v[2 * c + 1] = q.y;
      }
      nrm = 0.0;
#pragma unroll
      for (int c = 0; c < DP; ++c) nrm = fma(v[c], v[c], nrm);
#pragma unroll
      for (int c = 0; c < DP / 2; ++c)
        reinterpret_cast<double2 *>(dst2)[c] = make_double2(sc * v[2 * c], sc * v[2 * c + 1]);
      (is_row ? nx2 : ny2)[pidx] = nrm;
    }
    sal[tid] = a.alpha[pt];
#pragma unroll
    for (int e = 0; e < 4; ++e) tab[e * 256 + tid] = T.v[(e * 256 + tid) >> 4];   // consecutive lanes, consecutive words
  }
  __syncthreads();

  const int ibase = wn * 32;
  const double *bp = Brow + (ibase + g) * LDK + t;
  const double *bp2 = Brow2 + (ibase + g) * LDK + t;
  const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab) + ((unsigned)(lane & 15) << 3);
  // this thread's rows: ibase + nb * 8 + 2 t (+1), nb = 0..3
  double ali[4][2];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const double2 q = *reinterpret_cast<const double2 *>(sal + ibase + nb * 8 + 2 * t);
    ali[nb][0] = q.x;
    ali[nb][1] = q.y;
  }
  const double *kin = a.Kinv + (m0 + ibase + 2 * t) + (n0 + g) * a.ld;

  double s[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) s[q] = 0.0;

#pragma unroll 1
  for (int chunk = 0; chunk < 4; ++chunk) {
    const int jbase = chunk * 32 + wm * 16;
    double2 kv[2][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb)
        kv[mb][nb] = *reinterpret_cast<const double2 *>(kin + (int64_t)(jbase + mb * 8) * a.ld + nb * 8);
    const double *ap = Acol + (jbase + g) * LDK + t;
    const double *ap2 = Acol2 + (jbase + g) * LDK + t;
    double acc[2][4][2];
    double ac2[ARD ? 2 : 1][4][2];
    {
      const double y0 = ny[jbase + g], y1 = ny[jbase + 8 + g];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const double2 xn = *reinterpret_cast<const double2 *>(nx + ibase + nb * 8 + 2 * t);
        acc[0][nb][0] = y0 + xn.x;  acc[0][nb][1] = y0 + xn.y;
        acc[1][nb][0] = y1 + xn.x;  acc[1][nb][1] = y1 + xn.y;
      }
      if (ARD) {
        const double z0 = ny2[jbase + g], z1 = ny2[jbase + 8 + g];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
          const double2 xn = *reinterpret_cast<const double2 *>(nx2 + ibase + nb * 8 + 2 * t);
          ac2[0][nb][0] = z0 + xn.x;  ac2[0][nb][1] = z0 + xn.y;
          ac2[ARD ? 1 : 0][nb][0] = z1 + xn.x;  ac2[ARD ? 1 : 0][nb][1] = z1 + xn.y;
        }
      }
    }
#pragma unroll
    for (int kk = 0; kk < DP / 4; ++kk) {
      double af[2], bf[4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) af[mb] = ap[mb * 8 * LDK + kk * 4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) bf[nb] = bp[nb * 8 * LDK + kk * 4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(acc[mb][nb][0]), "+d"(acc[mb][nb][1])
                       : "d"(af[mb]), "d"(bf[nb]));
      if (ARD) {
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) af[mb] = ap2[mb * 8 * LDK + kk * 4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) bf[nb] = bp2[nb * 8 * LDK + kk * 4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(ac2[ARD ? mb : 0][nb][0]), "+d"(ac2[ARD ? mb : 0][nb][1])
                         : "d"(af[mb]), "d"(bf[nb]));
      }
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const double alj = sal[GT + jbase + mb * 8 + g];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double x[4], e[4], m[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int nb = 2 * q + (u >> 1);
          const double kvu = (u & 1) ? kv[mb][nb].y : kv[mb][nb].x;
          m[u] = fma(ali[nb][u & 1], alj, -kvu);
        }
        if (MODE == M_MATERN) {
          double r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double av = acc[mb][2 * q + (u >> 1)][u & 1];
            const double v = __hiloint2double(max(__double2hiint(av), 0x03c00000), __double2loint(av));
            const double y = rsqrt_seed(v);
            double gg = v * y;
            const double hh = 0.5 * y;
            const double ee = fma(-gg, hh, 0.5);
            gg = fma(gg, ee, gg);
            const double dd = fma(-gg, gg, v);
            r[u] = fma(dd, hh, gg);
            x[u] = -r[u];
          }
          lean_exp_v<4>(x, lc, tab_s, e);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            double p = P3 ? fma(f.q0, r[u], f.q1) : f.q1;
            p = fma(p, r[u], f.q2);
            p = fma(p, r[u], f.q3);
            s[0] = fma(m[u] * e[u], p * r[u], s[0]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = -acc[mb][2 * q + (u >> 1)][u & 1];
          lean_exp_v<4>(x, lc, tab_s, e);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double me = m[u] * e[u];
            if (MODE == M_SE) {
              s[0] = fma(me, acc[mb][2 * q + (u >> 1)][u & 1], s[0]);
              s[1] += me;
            } else {
              s[0] += me;
              s[1] = fma(me, ac2[ARD ? mb : 0][2 * q + (u >> 1)][u & 1], s[1]);
            }
          }
        }
      }
    }
  }
  if (MODE == M_SE) s[0] *= f.s0_scale;

#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = 2.0 * s[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NS + q] = v;
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
    for (int wdx = 0; wdx < GTHREADS / 32; ++wdx) v += red[wdx * NS + tid];
    a.partials[tlin * NS + tid] = v;
  }
}

// Any feature count (dp > 32): same tiling, row points re-read from global/L1, column points in
// shared memory, runtime loop over the features.  ARD exact mode (one sum per dimension) is not
// offered here.
template <int MODE>
__global__ void __launch_bounds__(GTHREADS) grad_trace_generic_kernel(CovParams cp, GradArgs a) {
  constexpr int NS = NSums<MODE, 0>::value;
  extern __shared__ __align__(16) double sm[];
  constexpr bool ARD = (MODE == M_ARD_REF);
  const int dp = a.dp;
  double *sx = sm;
  double *sxs = sm + GT * dp;
  double *sal = sm + (ARD ? 2 : 1) * GT * dp;
  double *red = sal + GT;
  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;
  int64_t tlin = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
  while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
  while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
  const int tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  const int64_t m0 = (int64_t)ti * GT, n0 = (int64_t)tj * GT;
  for (int i = tid; i < GT * dp; i += GTHREADS) {
    sx[i] = a.X[n0 * dp + i];
    if (ARD) sxs[i] = a.Xs[n0 * dp + i];
  }
  if (tid < GT) sal[tid] = a.alpha[n0 + tid];
  const int64_t r0 = m0 + 2 * tx;
  const double *x0 = a.X + r0 * dp, *x1 = x0 + dp;
  const double *xs0 = ARD ? a.Xs + r0 * dp : x0, *xs1 = xs0 + dp;
  const double al0 = a.alpha[r0], al1 = a.alpha[r0 + 1];
  __syncthreads();
  double s[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) s[q] = 0.0;
  const double *kinv = a.Kinv + r0 + n0 * a.ld;
  for (int cc = 0; cc < GT / 4; ++cc) {
    const int j = ty + 4 * cc;
    const int64_t col = n0 + j;
    const double2 kv = *reinterpret_cast<const double2 *>(kinv + (int64_t)j * a.ld);
    const double aj = sal[j];
    double d0 = 0.0, d1 = 0.0, p0 = 0.0, p1 = 0.0;
    for (int c = 0; c < dp; ++c) {
      const double y = sx[j * dp + c];
      const double e0 = x0[c] - y, e1 = x1[c] - y;
      if (MODE == M_PERIODIC || MODE == M_LAPLACIAN) {
        d0 += fabs(e0);
        d1 += fabs(e1);
      } else if (ARD) {
        const double ys = sxs[j * dp + c];
        const double f0 = xs0[c] - ys, f1 = xs1[c] - ys;
        d0 = fma(f0, f0, d0);
        d1 = fma(f1, f1, d1);
        p0 = fma(e0, e0, p0);
        p1 = fma(e1, e1, p1);
      } else {
        d0 = fma(e0, e0, d0);
        d1 = fma(e1, e1, d1);
      }
    }
    double w0 = (r0 > col) ? 2.0 : (r0 == col ? 1.0 : 0.0);
    double w1 = (r0 + 1 > col) ? 2.0 : (r0 + 1 == col ? 1.0 : 0.0);
    if (r0 >= a.n || col >= a.n) w0 = 0.0;
    if (r0 + 1 >= a.n || col >= a.n) w1 = 0.0;
    if (w0 != 0.0) accumulate<MODE, 0>(cp, w0 * (al0 * aj - kv.x), d0, p0, nullptr, s);
    if (w1 != 0.0) accumulate<MODE, 0>(cp, w1 * (al1 * aj - kv.y), d1, p1, nullptr, s);
  }
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = s[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NS + q] = v;
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
    for (int wdx = 0; wdx < GTHREADS / 32; ++wdx) v += red[wdx * NS + tid];
    a.partials[(int64_t)blockIdx.x * NS + tid] = v;
  }
}

// ---- kernel algebra (DGP_KERNEL_COMPOSITE) -----------------------------------------------------------
// d k / d theta_{f,i} = (d k / d v_f) * d v_f / d theta_{f,i}  (AdditiveCovariance.scala:77-84,
// MultiplicativeCovariance.scala:68-77): every base kernel accumulates its own moments exactly as in the
// single-kernel pass, with the pair weight M_ij multiplied by the cofactor d k / d v_f.  Moments of
// factor f start at moment_off[f]; the last moment is the noise one, sum M_ij 1[|x_i - x_j| == 0].
constexpr int MAX_COMPOSITE_MOMENTS = 128;

__device__ __forceinline__ void accumulate_rt(int mode, const CovParams &cp, double m, double dist, double plain2,
                                              double *s) {
  switch (mode) {
    case M_SE: accumulate<M_SE, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_ARD_REF: accumulate<M_ARD_REF, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_MATERN: accumulate<M_MATERN, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_PERIODIC: accumulate<M_PERIODIC, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_CAUCHY: accumulate<M_CAUCHY, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_LAPLACIAN: accumulate<M_LAPLACIAN, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_RATQUAD: accumulate<M_RATQUAD, 0>(cp, m, dist, plain2, nullptr, s); break;
    case M_TSTUDENT: accumulate<M_TSTUDENT, 0>(cp, m, dist, plain2, nullptr, s); break;
    default: accumulate<M_DIRAC, 0>(cp, m, dist, plain2, nullptr, s); break;
  }
}

struct CompositeModes {
  int mode[MAXF];
};

// FBMKernel.scala:96-120: a^H ln a + b^H ln b - c^H ln c with a = |x|^2, b = |y|^2, c = |x-y|^2 (its NaN guards
// compare against Double.NaN and never fire, so 0 ln 0 = NaN reaches the sum); exact mode: the true derivative.
__device__ __forceinline__ double fbm_dhurst(const CovParams &cp, const PairGeom &q) {
  const double h = cp.c1;
  if (cp.grad_mode == DGP_GRAD_EXACT) {
    auto term = [h](double a) { return a > 0.0 ? pow(a, h) * log(a) : 0.0; };
    return 0.5 * (term(q.nx) + term(q.ny) - term(q.r2));
  }
  return pow(q.nx, h) * log(q.nx) + pow(q.ny, h) * log(q.ny) - pow(q.r2, h) * log(q.r2);
}

template <int DP, int FT>
__device__ __forceinline__ void composite_pair(const CompositeDev &c, const CompositeModes &md, const double *x,
                                               const double *y, int dp, double m, double *s) {
  const int n = DP > 0 ? DP : dp;
  PairGeom q;
  double ard[MAX_ARD_SLOTS] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < n; ++k) {
    const double e = x[k] - y[k];
    const double e2 = e * e;
    q.r2 += e2;
    if (c.any_l1) q.l1 += fabs(e);
    if (c.any_ip) {
      q.dot = fma(x[k], y[k], q.dot);
      q.nx = fma(x[k], x[k], q.nx);
      q.ny = fma(y[k], y[k], q.ny);
    }
#pragma unroll
    for (int a = 0; a < MAX_ARD_SLOTS; ++a)
      if (a < c.nslots && k < MAX_ARD_DIM) ard[a] = fma(c.w[a][k], c.slot_l1[a] ? fabs(e) : e2, ard[a]);
  }
  const double r2 = q.r2;
  double v[FT], cof[FT];
  auto dist_of = [&](int f) {
    const int id = c.metric[f];
    return id == 0 ? q.r2 : (id == 1 ? q.l1 : ard[id - 2]);
  };
  auto value_of = [&](int f) {
    return kind_is_inner_product(c.cp[f].kind) ? cov_value_ip(c.cp[f], q) : cov_value_shared(c.cp[f], dist_of(f));
  };
  auto accumulate_factor = [&](int f) {
    const double dist = dist_of(f);
    const double w = m * cof[f];
    double *sf = s + c.moment_off[f];
    if (c.ard_exact[f]) {
      // exact dk/db_c = k (x_c - y_c)^2 / b_c^3: one moment per dimension (cf. M_ARD_EXACT)
      const double e = exp(-dist * c.cp[f].c1);
      sf[0] = fma(w, e, sf[0]);
      const double we = w * e;
      const int k0 = c.dim0[f], k1 = k0 + c.dimn[f];
#pragma unroll
      for (int k = 0; k < n; ++k) {  // static index into x (a register array when DP > 0)
        if (k < k0 || k >= k1) continue;
        const double df = x[k] - y[k];
        sf[1 + k - k0] = fma(we, df * df, sf[1 + k - k0]);
      }
    } else if (md.mode[f] == M_FBM) {
      sf[0] = fma(w, fbm_dhurst(c.cp[f], q), sf[0]);
    } else if (md.mode[f] == M_POLY) {
      // no derivative in the reference (stub returning 0.0 per hyper-parameter)
    } else {
      accumulate_rt(md.mode[f], c.cp[f], w, dist, r2, sf);
    }
  };
  if (FactorLoop<FT>::unrolled) {
#pragma unroll
    for (int f = 0; f < FT; ++f) v[f] = (f < c.F) ? value_of(f) : 1.0;
    composite_cofactors<FT>(c, v, cof);
#pragma unroll
    for (int f = 0; f < FT; ++f)
      if (f < c.F) accumulate_factor(f);
  } else {
    for (int f = 0; f < c.F; ++f) v[f] = value_of(f);
    composite_cofactors<FT>(c, v, cof);
    for (int f = 0; f < c.F; ++f) accumulate_factor(f);
  }
  if (r2 == 0.0) s[c.moment_off[c.F]] += m;
}

template <int DP, int FT>
__global__ void __launch_bounds__(GTHREADS) grad_trace_composite_kernel(CompositeDev c, CompositeModes md, GradArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int dp = DP > 0 ? DP : a.dp;
  const int NS = c.moment_off[c.F] + 1;
  double *sx = sm;
  double *sal = sm + GT * dp;
  double *red = sal + GT;
  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;
  int64_t tlin = blockIdx.x;
  int ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
  while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
  while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
  const int tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  const int64_t m0 = (int64_t)ti * GT, n0 = (int64_t)tj * GT;
  for (int i = tid; i < GT * dp; i += GTHREADS) sx[i] = a.X[n0 * dp + i];
  if (tid < GT) sal[tid] = a.alpha[n0 + tid];
  const int64_t r0 = m0 + 2 * tx;
  const double *p0 = a.X + r0 * dp, *p1 = p0 + dp;
  double x0[DP > 0 ? DP : 1], x1[DP > 0 ? DP : 1];
  if (DP > 0) {
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      x0[k] = p0[k];
      x1[k] = p1[k];
    }
  }
  const double al0 = a.alpha[r0], al1 = a.alpha[r0 + 1];
  __syncthreads();
  double s[MAX_COMPOSITE_MOMENTS];
  for (int q = 0; q < NS; ++q) s[q] = 0.0;
  const double *kinv = a.Kinv + r0 + n0 * a.ld;
  for (int cc = 0; cc < GT / 4; ++cc) {
    const int j = ty + 4 * cc;
    const int64_t col = n0 + j;
    const double2 kv = *reinterpret_cast<const double2 *>(kinv + (int64_t)j * a.ld);
    const double aj = sal[j];
    double w0 = (r0 > col) ? 2.0 : (r0 == col ? 1.0 : 0.0);
    double w1 = (r0 + 1 > col) ? 2.0 : (r0 + 1 == col ? 1.0 : 0.0);
    if (r0 >= a.n || col >= a.n) w0 = 0.0;
    if (r0 + 1 >= a.n || col >= a.n) w1 = 0.0;
    if (w0 != 0.0) composite_pair<DP, FT>(c, md, DP > 0 ? x0 : p0, sx + j * dp, dp, w0 * (al0 * aj - kv.x), s);
    if (w1 != 0.0) composite_pair<DP, FT>(c, md, DP > 0 ? x1 : p1, sx + j * dp, dp, w1 * (al1 * aj - kv.y), s);
  }
  const int warp = tid >> 5, lane = tid & 31;
  for (int q = 0; q < NS; ++q) {
    double v = s[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * NS + q] = v;
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
    for (int wdx = 0; wdx < GTHREADS / 32; ++wdx) v += red[wdx * NS + tid];
    a.partials[(int64_t)blockIdx.x * NS + tid] = v;
  }
}

template <int DP, int FT>
int32_t launch_composite_trace(Ctx *ctx, cudaStream_t st, const CompositeDev &c, const CompositeModes &md,
                               const GradArgs &a, int64_t tiles, int NS) {
  size_t smem = ((size_t)GT * a.dp + GT + (GTHREADS / 32) * NS) * sizeof(double);
  if (smem > 200 * 1024) return ctx->fail(DGP_ERR_SHAPE, "gradEnergy: too many features (dp=%d)", a.dp);
  if (smem > 48 * 1024)
    DGP_CUDA_OK(ctx, cudaFuncSetAttribute(grad_trace_composite_kernel<DP, FT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  grad_trace_composite_kernel<DP, FT><<<(unsigned)tiles, GTHREADS, smem, st>>>(c, md, a);
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

template <int FT>
int32_t launch_composite_trace_dp(Ctx *ctx, cudaStream_t st, const CompositeDev &c, const CompositeModes &md,
                                  const GradArgs &a, int64_t tiles, int NS) {
  if (!FactorLoop<FT>::unrolled) return launch_composite_trace<0, FT>(ctx, st, c, md, a, tiles, NS);
  switch (a.dp) {
    case 4: return launch_composite_trace<4, FT>(ctx, st, c, md, a, tiles, NS);
    case 8: return launch_composite_trace<8, FT>(ctx, st, c, md, a, tiles, NS);
    case 12: return launch_composite_trace<12, FT>(ctx, st, c, md, a, tiles, NS);
    case 16: return launch_composite_trace<16, FT>(ctx, st, c, md, a, tiles, NS);
    case 20: return launch_composite_trace<20, FT>(ctx, st, c, md, a, tiles, NS);
    case 24: return launch_composite_trace<24, FT>(ctx, st, c, md, a, tiles, NS);
    case 32: return launch_composite_trace<32, FT>(ctx, st, c, md, a, tiles, NS);
    default: return launch_composite_trace<0, FT>(ctx, st, c, md, a, tiles, NS);
  }
}

template <int MODE>
int32_t launch_generic(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int64_t tiles, int *ns_out) {
  constexpr int NS = NSums<MODE, 0>::value;
  constexpr bool ARD = (MODE == M_ARD_REF);
  size_t smem = ((size_t)(ARD ? 2 : 1) * GT * a.dp + GT + (GTHREADS / 32) * NS) * sizeof(double);
  if (smem > 200 * 1024) return ctx->fail(DGP_ERR_SHAPE, "gradEnergy: too many features (dp=%d)", a.dp);
  if (smem > 48 * 1024)
    DGP_CUDA_OK(ctx, cudaFuncSetAttribute(grad_trace_generic_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
  grad_trace_generic_kernel<MODE><<<(unsigned)tiles, GTHREADS, smem, st>>>(cp, a);
  DGP_CUDA_OK(ctx, cudaGetLastError());
  reduce_partials_kernel<<<NS, 1024, 0, st>>>(a.partials, tiles, NS, a.sums);
  ctx->launches += 2;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  *ns_out = NS;
  return DGP_OK;
}

template <int MODE, int DP>
int32_t launch(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int64_t tiles, int *ns_out) {
  constexpr int NS = NSums<MODE, DP>::value;
  constexpr bool ARD = (MODE == M_ARD_REF || MODE == M_ARD_EXACT);
  size_t smem = ((size_t)(ARD ? 2 : 1) * GT * DP + GT + (GTHREADS / 32) * NS) * sizeof(double);
  if (smem > 48 * 1024)
    DGP_CUDA_OK(ctx, cudaFuncSetAttribute(grad_trace_kernel<MODE, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
  grad_trace_kernel<MODE, DP><<<(unsigned)tiles, GTHREADS, smem, st>>>(cp, a);
  DGP_CUDA_OK(ctx, cudaGetLastError());
  reduce_partials_kernel<<<NS, 1024, 0, st>>>(a.partials, tiles, NS, a.sums);
  ctx->launches += 2;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  *ns_out = NS;
  return DGP_OK;
}

// Lean pass over the strictly-lower tiles + exact-difference pass over the diagonal / padded tiles + reduction.
template <int MODE, int DP>
int32_t launch_lean_trace(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int64_t tiles, int *ns_out) {
  constexpr int NS = NSums<MODE, DP>::value;
  constexpr bool ARD = (MODE == M_ARD_REF);
  const int64_t mt = a.np / GT;
  LeanTraceCov f;
  f.scale = (MODE == M_MATERN) ? cp.c1 : sqrt(cp.c1);
  f.q0 = f.q1 = f.q2 = f.q3 = 0.0;
  f.s0_scale = (MODE == M_SE) ? 1.0 / cp.c1 : 1.0;
  if (MODE == M_MATERN) {
    // with rho = c1 r and u = c2 r = kappa rho:  (c1 r sum(u) - q(u)) / c3 = rho * sum_k A_k rho^k,
    //   A_k = (coef[p-k] kappa^k - (k+1) coef[p-k-1] kappa^(k+1)) / c3,  k = 0..p  (second term absent for k = p)
    const double kappa = cp.c2 / cp.c1;
    double *q[4] = {&f.q0, &f.q1, &f.q2, &f.q3};
    for (int k = 0; k <= cp.p; ++k) {
      double v = cp.coef[cp.p - k] * pow(kappa, (double)k);
      if (k + 1 <= cp.p) v -= (double)(k + 1) * cp.coef[cp.p - k - 1] * pow(kappa, (double)(k + 1));
      *q[3 - k] = v / cp.c3;
    }
  }
  const ExpTab T = lean_exp_table(1.0);
  const LeanConst lc = lean_constants();
  const size_t lsmem = ((size_t)64 * 16 + (ARD ? 2 : 1) * (2 * GT * lean_ldk<DP>() + 2 * GT) + 2 * GT + 8 * NS + 2) * sizeof(double);
  const size_t esmem = ((size_t)(ARD ? 2 : 1) * GT * DP + GT + (GTHREADS / 32) * NS) * sizeof(double);
  auto go = [&](auto kern) -> int32_t {
    if (lsmem > 48 * 1024)
      DGP_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
    kern<<<(unsigned)(mt * (mt - 1) / 2), GTHREADS, lsmem, st>>>(f, lc, T, a);
    return DGP_OK;
  };
  if (MODE == M_MATERN && cp.p == 3) DGP_TRY(go(grad_trace_lean_kernel<MODE, DP, (MODE == M_MATERN)>));
  else DGP_TRY(go(grad_trace_lean_kernel<MODE, DP, false>));
  DGP_CUDA_OK(ctx, cudaGetLastError());
  GradArgs e = a;
  e.edge_only = 1;
  if (esmem > 48 * 1024)
    DGP_CUDA_OK(ctx, cudaFuncSetAttribute(grad_trace_kernel<MODE, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)esmem));
  const int64_t etiles = mt + ((a.np != a.n) ? mt - 1 : 0);
  grad_trace_kernel<MODE, DP><<<(unsigned)etiles, GTHREADS, esmem, st>>>(cp, e);
  DGP_CUDA_OK(ctx, cudaGetLastError());
  reduce_partials_kernel<<<NS, 1024, 0, st>>>(a.partials, tiles, NS, a.sums);
  ctx->launches += 3;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  *ns_out = NS;
  return DGP_OK;
}

template <int MODE>
int32_t launch_lean_dp(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int64_t tiles, int *ns) {
  switch (a.dp) {
    case 4: return launch_lean_trace<MODE, 4>(ctx, st, cp, a, tiles, ns);
    case 8: return launch_lean_trace<MODE, 8>(ctx, st, cp, a, tiles, ns);
    case 12: return launch_lean_trace<MODE, 12>(ctx, st, cp, a, tiles, ns);
    case 16: return launch_lean_trace<MODE, 16>(ctx, st, cp, a, tiles, ns);
    case 20: return launch_lean_trace<MODE, 20>(ctx, st, cp, a, tiles, ns);
    case 24: return launch_lean_trace<MODE, 24>(ctx, st, cp, a, tiles, ns);
    case 32: return launch_lean_trace<MODE, 32>(ctx, st, cp, a, tiles, ns);
    default: return DGP_ERR_SHAPE;   // not reached: lean_trace_ok() checks the feature count
  }
}

bool lean_trace_ok(const CovParams &cp, const GradArgs &a, int mode) {
  if (!a.lean || a.np < 4096 || a.np - a.n >= GT) return false;   // below ~4096 points the third launch costs more than it saves
  if (a.dp != 4 && a.dp != 8 && a.dp != 12 && a.dp != 16 && a.dp != 20 && a.dp != 24 && a.dp != 32) return false;
  if (mode == M_MATERN) return cp.p >= 1 && cp.p <= 3 && cp.c1 > 0.0 && std::isfinite(cp.c2 / cp.c1);
  return (mode == M_SE || mode == M_ARD_REF) && cp.c1 > 0.0 && std::isfinite(1.0 / cp.c1) && std::isfinite(sqrt(cp.c1));
}

template <int MODE>
int32_t launch_dp(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int64_t tiles, int *ns) {
  switch (a.dp) {
    case 4: return launch<MODE, 4>(ctx, st, cp, a, tiles, ns);
    case 8: return launch<MODE, 8>(ctx, st, cp, a, tiles, ns);
    case 12: return launch<MODE, 12>(ctx, st, cp, a, tiles, ns);
    case 16: return launch<MODE, 16>(ctx, st, cp, a, tiles, ns);
    case 20: return launch<MODE, 20>(ctx, st, cp, a, tiles, ns);
    case 24: return launch<MODE, 24>(ctx, st, cp, a, tiles, ns);
    case 32: return launch<MODE, 32>(ctx, st, cp, a, tiles, ns);
    default:
      if (MODE == M_ARD_EXACT)
        return ctx->fail(DGP_ERR_SHAPE, "exact ARD gradients support up to 32 features, got dp=%d", a.dp);
      return launch_generic<(MODE == M_ARD_EXACT ? M_ARD_REF : MODE)>(ctx, st, cp, a, tiles, ns);
  }
}

int mode_of(const CovParams &cp) {
  switch (cp.kind) {
    case DGP_KERNEL_SE:
    case DGP_KERNEL_RBF: return M_SE;
    case DGP_KERNEL_ARD_SE: return cp.grad_mode == DGP_GRAD_EXACT ? M_ARD_EXACT : M_ARD_REF;
    case DGP_KERNEL_MATERN: return M_MATERN;
    case DGP_KERNEL_PERIODIC: return M_PERIODIC;
    case DGP_KERNEL_CAUCHY: return M_CAUCHY;
    case DGP_KERNEL_LAPLACIAN: return M_LAPLACIAN;
    case DGP_KERNEL_RATQUAD: return M_RATQUAD;
    case DGP_KERNEL_TSTUDENT: return M_TSTUDENT;
    case DGP_KERNEL_POLYNOMIAL: return M_POLY;
    case DGP_KERNEL_FBM: return M_FBM;
    default: return M_DIRAC;
  }
}

// ---- per-pair moments: the gradient MATRICES (SVMKernel.buildKernelGradMatrix / buildCrossKernelGradMatrix) ----------
// dK_ij / dtheta is what the trace pass contributes for the weight matrix M = e_i e_j', so one thread per pair runs the
// very same accumulate<> / composite_pair<> with m = 1 into its own moment vector; grad_finish (linear in the moments)
// then maps each pair's moments to the reference's derivatives on the host.  Not a hot path: runtime feature loop.
constexpr int PAIR_NS_MAX = MAX_COMPOSITE_MOMENTS;

__global__ void __launch_bounds__(128) pair_moments_kernel(CovParams cp, int mode, const double *X1, const double *X1s,
                                                           const double *X2, const double *X2s, int64_t n1, int64_t c0,
                                                           int64_t ncols, int dp, int ns, double *out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n1 * ncols) return;
  const int64_t i = idx % n1, jl = idx / n1, j = c0 + jl;
  double s[MAX_ARD_DIM + 2 > 8 ? MAX_ARD_DIM + 2 : 8];
  for (int q = 0; q < ns; ++q) s[q] = 0.0;
  const double *x = X1 + i * dp, *y = X2 + j * dp;
  double d = 0.0, p = 0.0, sq[MAX_ARD_DIM];
  for (int c = 0; c < MAX_ARD_DIM; ++c) sq[c] = 0.0;
  const bool l1 = mode == M_PERIODIC || mode == M_LAPLACIAN, ard = mode == M_ARD_REF || mode == M_ARD_EXACT;
  for (int c = 0; c < dp; ++c) {
    const double e = x[c] - y[c];
    if (l1) d += fabs(e);
    else if (ard) {
      const double f = X1s[i * dp + c] - X2s[j * dp + c];
      d = fma(f, f, d);
      p = fma(e, e, p);
      if (c < MAX_ARD_DIM) sq[c] = e * e;
    } else d = fma(e, e, d);
  }
  switch (mode) {
    case M_SE: accumulate<M_SE, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_ARD_REF: accumulate<M_ARD_REF, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_ARD_EXACT: accumulate<M_ARD_EXACT, MAX_ARD_DIM>(cp, 1.0, d, p, sq, s); break;
    case M_MATERN: accumulate<M_MATERN, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_PERIODIC: accumulate<M_PERIODIC, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_CAUCHY: accumulate<M_CAUCHY, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_LAPLACIAN: accumulate<M_LAPLACIAN, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_RATQUAD: accumulate<M_RATQUAD, 0>(cp, 1.0, d, p, nullptr, s); break;
    case M_TSTUDENT: accumulate<M_TSTUDENT, 0>(cp, 1.0, d, p, nullptr, s); break;
    default: accumulate<M_DIRAC, 0>(cp, 1.0, d, p, nullptr, s); break;
  }
  double *o = out + (i + jl * n1) * ns;
  for (int q = 0; q < ns; ++q) o[q] = s[q];
}

__global__ void __launch_bounds__(64) pair_moments_composite_kernel(CompositeDev c, CompositeModes md, const double *X1,
                                                                    const double *X2, int64_t n1, int64_t c0,
                                                                    int64_t ncols, int dp, int ns, double *out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n1 * ncols) return;
  const int64_t i = idx % n1, jl = idx / n1, j = c0 + jl;
  double s[PAIR_NS_MAX];
  for (int q = 0; q < ns; ++q) s[q] = 0.0;
  composite_pair<0, MAXF>(c, md, X1 + i * dp, X2 + j * dp, dp, 1.0, s);
  double *o = out + (i + jl * n1) * ns;
  for (int q = 0; q < ns; ++q) o[q] = s[q];
}

}  // namespace

int32_t grad_pair_moments(Ctx *ctx, cudaStream_t st, const CovParams &cp, const double *X1, const double *X1s,
                          const double *X2, const double *X2s, int64_t n1, int64_t c0, int64_t ncols, int dp, double *out,
                          int *ns_out) {
  const int64_t pairs = n1 * ncols;
  if (pairs <= 0) return DGP_OK;
  if (cp.kind == DGP_KERNEL_COMPOSITE) {
    if (!cp.comp) return ctx->fail(DGP_ERR_ARG, "grad matrix: composite kernel without an expression");
    const CompositeDev &c = *cp.comp;
    const int NS = c.moment_off[c.F] + 1;
    if (NS > PAIR_NS_MAX) return ctx->fail(DGP_ERR_SHAPE, "composite gradient needs %d moments, at most %d", NS, PAIR_NS_MAX);
    CompositeModes md;
    for (int f = 0; f < MAXF; ++f) md.mode[f] = f < c.F ? mode_of(c.cp[f]) : M_DIRAC;
    pair_moments_composite_kernel<<<(unsigned)((pairs + 63) / 64), 64, 0, st>>>(c, md, X1, X2, n1, c0, ncols, dp, NS, out);
    *ns_out = NS;
  } else {
    const int mode = mode_of(cp);
    if (mode == M_ARD_EXACT && dp > MAX_ARD_DIM)
      return ctx->fail(DGP_ERR_SHAPE, "exact ARD gradients support up to %d features, got dp=%d", MAX_ARD_DIM, dp);
    const int NS = mode == M_ARD_EXACT ? MAX_ARD_DIM + 2 : grad_partials_per_tile(cp, dp);
    pair_moments_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(cp, mode, X1, X1s, X2, X2s, n1, c0, ncols, dp, NS,
                                                                         out);
    *ns_out = NS;
  }
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int grad_partials_per_tile(const CovParams &cp, int dp) {
  if (cp.kind == DGP_KERNEL_COMPOSITE) return cp.comp ? cp.comp->moment_off[cp.comp->F] + 1 : 1;
  switch (mode_of(cp)) {
    case M_ARD_EXACT: return dp + 2;
    case M_MATERN:
    case M_DIRAC:
    case M_CAUCHY:
    case M_LAPLACIAN:
    case M_TSTUDENT: return 2;
    default: return 3;
  }
}

int32_t grad_trace(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int *ns) {
  if (a.np % GT) return ctx->fail(DGP_ERR_SHAPE, "grad_trace: padded n must be a multiple of 128");
  const int64_t mt = a.np / GT, tiles = mt * (mt + 1) / 2;
  if (cp.kind == DGP_KERNEL_COMPOSITE) {
    if (!cp.comp) return ctx->fail(DGP_ERR_ARG, "grad_trace: composite kernel without an expression");
    const CompositeDev &c = *cp.comp;
    const int NS = c.moment_off[c.F] + 1;
    if (NS > MAX_COMPOSITE_MOMENTS)
      return ctx->fail(DGP_ERR_SHAPE, "composite gradient needs %d moments, at most %d are supported", NS,
                       MAX_COMPOSITE_MOMENTS);
    CompositeModes md;
    for (int f = 0; f < MAXF; ++f) md.mode[f] = f < c.F ? mode_of(c.cp[f]) : M_DIRAC;
    if (c.F <= 4) DGP_TRY(launch_composite_trace_dp<4>(ctx, st, c, md, a, tiles, NS));
    else DGP_TRY(launch_composite_trace_dp<MAXF>(ctx, st, c, md, a, tiles, NS));
    reduce_partials_kernel<<<NS, 1024, 0, st>>>(a.partials, tiles, NS, a.sums);
    ctx->launches += 2;
    DGP_CUDA_OK(ctx, cudaGetLastError());
    *ns = NS;
    return DGP_OK;
  }
  const int mode = mode_of(cp);
  if (lean_trace_ok(cp, a, mode)) {
    if (mode == M_SE) return launch_lean_dp<M_SE>(ctx, st, cp, a, tiles, ns);
    if (mode == M_ARD_REF) return launch_lean_dp<M_ARD_REF>(ctx, st, cp, a, tiles, ns);
    return launch_lean_dp<M_MATERN>(ctx, st, cp, a, tiles, ns);
  }
  switch (mode) {
    case M_SE: return launch_dp<M_SE>(ctx, st, cp, a, tiles, ns);
    case M_ARD_REF: return launch_dp<M_ARD_REF>(ctx, st, cp, a, tiles, ns);
    case M_ARD_EXACT: return launch_dp<M_ARD_EXACT>(ctx, st, cp, a, tiles, ns);
    case M_MATERN: return launch_dp<M_MATERN>(ctx, st, cp, a, tiles, ns);
    case M_PERIODIC: return launch_dp<M_PERIODIC>(ctx, st, cp, a, tiles, ns);
    case M_CAUCHY: return launch_dp<M_CAUCHY>(ctx, st, cp, a, tiles, ns);
    case M_LAPLACIAN: return launch_dp<M_LAPLACIAN>(ctx, st, cp, a, tiles, ns);
    case M_RATQUAD: return launch_dp<M_RATQUAD>(ctx, st, cp, a, tiles, ns);
    case M_TSTUDENT: return launch_dp<M_TSTUDENT>(ctx, st, cp, a, tiles, ns);
    default: return launch_dp<M_DIRAC>(ctx, st, cp, a, tiles, ns);
  }
}

// Map the moments to the reference's derivatives, in the order [hyp..., noise].
void grad_finish(const Spec &sp, int d, const double *S, int ns, double *g) {
  const double qnan = nan("");
  if (sp.kind == DGP_KERNEL_COMPOSITE) {
    const Composite &cm = *sp.comp;
    int pos = 0;
    for (int f = 0; f < cm.dev.F; ++f) {
      double tmp[DGP_MAX_HYP + 1];
      grad_finish(cm.factor[f], cm.factor_d[f], S + cm.dev.moment_off[f],
                  cm.dev.moment_off[f + 1] - cm.dev.moment_off[f], tmp);
      for (int i = 0; i < cm.factor[f].n_hyp; ++i) g[pos++] = tmp[i];
    }
    g[pos] = (sp.noise == 0.0) ? qnan : S[ns - 1];
    return;
  }
  const double noise_m = S[ns - 1];
  switch (sp.kind) {
    case DGP_KERNEL_SE: {
      double l = sp.hyp[0], a = sp.hyp[1];
      g[0] = a * a * S[0] / (fabs(l) * fabs(l) * fabs(l));  // k |x-y|^2 / |l|^3, k = a^2 e
      g[1] = 2.0 * a * S[1];
      break;
    }
    case DGP_KERNEL_RBF: {
      double l = sp.hyp[0];
      g[0] = S[0] / (fabs(l) * fabs(l) * fabs(l));
      break;
    }
    case DGP_KERNEL_ARD_SE: {
      double h = sp.hyp[0];
      g[0] = 2.0 * (h * h * S[0]) / h;  // 2 k / h
      for (int c = 0; c < d; ++c) {
        double b = sp.hyp[1 + c];
        double m = (sp.grad_mode == DGP_GRAD_EXACT) ? S[1 + c] : S[1];
        g[1 + c] = h * h * m / (b * b * b);  // math.pow(b, 3.0): signed, as the reference
      }
      break;
    }
    case DGP_KERNEL_MATERN:
      g[0] = qnan;  // "p" has no derivative in the reference (blocked)
      g[1] = S[0];
      break;
    case DGP_KERNEL_PERIODIC: {
      double sg = sp.hyp[0], l = sp.hyp[1], f = sp.hyp[2];
      double s2 = sg * sg;
      g[0] = 2.0 * S[0] / sg;
      g[1] = 4.0 * s2 * S[1] / l;
      g[2] = -2.0 * s2 * S[1] / f;
      break;
    }
    case DGP_KERNEL_DIRAC:
      g[0] = (sp.hyp[0] == 0.0) ? qnan : S[0];  // |s| 1[..] / |s|
      break;
    case DGP_KERNEL_CAUCHY: {
      double sg = sp.hyp[0];
      g[0] = 2.0 * S[0] / (sg * sg * sg);
      break;
    }
    case DGP_KERNEL_LAPLACIAN:
      g[0] = S[0] / (sp.hyp[0] * sp.hyp[0]);
      break;
    case DGP_KERNEL_RATQUAD: {
      double mu = sp.hyp[0], l = sp.hyp[1];
      double e = -0.5 * ((double)d + mu);
      g[0] = S[1];
      g[1] = -2.0 * e * S[0] / (mu * l * l * l);
      break;
    }
    case DGP_KERNEL_TSTUDENT:
      g[0] = S[0];
      break;
    case DGP_KERNEL_POLYNOMIAL:
      g[0] = 0.0;  // LocalSVMKernel.scala:37-41: "gradientAt function not defined, continuing with a stub"
      g[1] = 0.0;
      break;
    case DGP_KERNEL_FBM:
      g[0] = S[0];
      break;
  }
  g[sp.n_hyp] = (sp.noise == 0.0) ? qnan : noise_m;
}

}  // namespace dgp
