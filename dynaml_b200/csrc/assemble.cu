// assemble.cu -- kernel-matrix assembly K = k(X1, X2) (+ Dirac noise) in FP64.
//
// Replaces SVMKernel.buildSVMKernelMatrix / crossKernelMatrix / buildPartitionedKernelMatrix
// (CORE/kernels/SVMKernel.scala:71-105, 181-248), which evaluate one boxed closure per pair.
//
// One CTA produces a 128x128 tile.  Thread (tx, ty) owns rows {2tx, 2tx+1} (held in registers)
// and columns ty, ty+4, ...; the column points sit in shared memory and are read as warp-wide
// broadcasts.  Distances are exact differences (sum (x_c - y_c)^2 or sum |x_c - y_c|), so the
// diagonal and duplicated points give dist == 0 exactly -- the condition DiracKernel tests
// (DiracKernel.scala:43-47).  Every store is a 16-byte vector store and a warp writes 512
// contiguous bytes of one column; the kernel is bound by HBM write bandwidth or by the FP64
// pipe (exp/sqrt/sin), whichever ncu shows.
#include <cmath>

#include "assemble.cuh"
#include "lean.cuh"

namespace dgp {

namespace {

constexpr int AT = 128;      // tile edge
constexpr int ATHREADS = 256;

template <int KIND, int DP>
__global__ void __launch_bounds__(ATHREADS) assemble_kernel(CovParams cp, AssembleArgs a) {
  extern __shared__ __align__(16) double sm[];  // column points: AT x dp, row-major
  const int dp = (DP > 0) ? DP : a.dp;
  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;

  int64_t tlin = blockIdx.x;
  int ti, tj;
  if (a.lower_only) {
    ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
    while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
    tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  } else {
    int mt = (int)(a.rows_p / AT);
    tj = (int)(tlin / mt);
    ti = (int)(tlin % mt);
  }
  int64_t m0 = (int64_t)ti * AT + a.row_off, n0 = (int64_t)tj * AT + a.col_off;
  const int64_t lm0 = m0, ln0 = n0;  // position inside `out`
  if (a.bc_T > 0) {
    m0 = ((int64_t)(lm0 / a.bc_T + a.bc_row0) * a.bc_Pr + a.bc_pr) * a.bc_T + lm0 % a.bc_T;
    n0 = ((int64_t)(ln0 / a.bc_T + a.bc_col0) * a.bc_Pc + a.bc_pc) * a.bc_T + ln0 % a.bc_T;
    if (n0 > m0 + AT - 1) return;  // above the global diagonal (uniform, before any barrier)
  }

  // column points -> shared (contiguous AT*dp doubles in global memory)
  {
    const double2 *src = reinterpret_cast<const double2 *>(a.X2 + n0 * dp);
    double2 *dst = reinterpret_cast<double2 *>(sm);
    for (int i = tid; i < AT * dp / 2; i += ATHREADS) dst[i] = src[i];
  }

  const int64_t r0 = m0 + 2 * tx;
  const bool sym = a.symmetric != 0;
  double *out = a.out + (lm0 - a.row_off + 2 * tx) + (ln0 - a.col_off) * a.ld;

  if (DP > 0) {
    double x0[DP > 0 ? DP : 1], x1[DP > 0 ? DP : 1];
    {
      const double2 *p0 = reinterpret_cast<const double2 *>(a.X1 + r0 * DP);
#pragma unroll
      for (int c = 0; c < DP / 2; ++c) {
        double2 v = p0[c];
        x0[2 * c] = v.x;
        x0[2 * c + 1] = v.y;
      }
#pragma unroll
      for (int c = 0; c < DP / 2; ++c) {
        double2 v = p0[DP / 2 + c];
        x1[2 * c] = v.x;
        x1[2 * c + 1] = v.y;
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int cc = 0; cc < AT / 4; ++cc) {
      const int j = ty + 4 * cc;
      const double2 *y = reinterpret_cast<const double2 *>(sm + j * DP);
      double d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int c = 0; c < DP / 2; ++c) {
        double2 v = y[c];
        double e0 = x0[2 * c] - v.x, e1 = x0[2 * c + 1] - v.y;
        double f0 = x1[2 * c] - v.x, f1 = x1[2 * c + 1] - v.y;
        if (kind_is_l1(KIND)) {
          d0 += fabs(e0) + fabs(e1);
          d1 += fabs(f0) + fabs(f1);
        } else {
          d0 = fma(e0, e0, d0);
          d0 = fma(e1, e1, d0);
          d1 = fma(f0, f0, d1);
          d1 = fma(f1, f1, d1);
        }
      }
      double k0 = cov_value<KIND>(cp, d0), k1 = cov_value<KIND>(cp, d1);
      if (d0 == 0.0) k0 += cp.noise_abs;
      if (d1 == 0.0) k1 += cp.noise_abs;
      const int64_t col = n0 + j;
      // padding: identity for a symmetric matrix (keeps it PD, log-det unchanged), zero otherwise
      if (r0 >= a.rows || col >= a.cols) k0 = (sym && r0 == col) ? 1.0 : 0.0;
      if (r0 + 1 >= a.rows || col >= a.cols) k1 = (sym && r0 + 1 == col) ? 1.0 : 0.0;
      *reinterpret_cast<double2 *>(out + (int64_t)j * a.ld) = make_double2(k0, k1);
    }
  } else {
    // generic feature count: row points are re-read from global/L1 (cold path, d > 32)
    __syncthreads();
    const double *p0 = a.X1 + r0 * dp, *p1 = p0 + dp;
    for (int cc = 0; cc < AT / 4; ++cc) {
      const int j = ty + 4 * cc;
      const double *y = sm + j * dp;
      double d0 = 0.0, d1 = 0.0;
      for (int c = 0; c < dp; ++c) {
        double v = y[c];
        double e0 = p0[c] - v, f0 = p1[c] - v;
        if (kind_is_l1(KIND)) {
          d0 += fabs(e0);
          d1 += fabs(f0);
        } else {
          d0 = fma(e0, e0, d0);
          d1 = fma(f0, f0, d1);
        }
      }
      double k0 = cov_value<KIND>(cp, d0), k1 = cov_value<KIND>(cp, d1);
      if (d0 == 0.0) k0 += cp.noise_abs;
      if (d1 == 0.0) k1 += cp.noise_abs;
      const int64_t col = n0 + j;
      if (r0 >= a.rows || col >= a.cols) k0 = (sym && r0 == col) ? 1.0 : 0.0;
      if (r0 + 1 >= a.rows || col >= a.cols) k1 = (sym && r0 + 1 == col) ? 1.0 : 0.0;
      *reinterpret_cast<double2 *>(out + (int64_t)j * a.ld) = make_double2(k0, k1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel algebra (DGP_KERNEL_COMPOSITE): sums / products / scalings of base kernels evaluated per
// pair from one pass over the coordinate differences -- |x-y|_2^2, |x-y|_1 when some factor needs it,
// and a weighted |.|^2 per ARD factor -- then combined as sum_t coef_t prod_f k_f^e_tf.  Replaces the
// nested closures of AdditiveCovariance.evaluateAt / MultiplicativeCovariance.evaluateAt
// (AdditiveCovariance.scala:47-52, MultiplicativeCovariance.scala:33-39), which re-filter the
// hyper-parameter Map for every pair.  Same tiling as assemble_kernel's generic branch.
// ---------------------------------------------------------------------------------------------
struct PairMetrics {
  PairGeom g;
  double ard[MAX_ARD_SLOTS] = {0.0, 0.0, 0.0, 0.0};
};

// DP > 0: compile-time feature count (x is a register array, the loop unrolls); DP == 0: runtime count `dp`.
template <int DP>
__device__ __forceinline__ void pair_metrics(const CompositeDev &c, const double *x, const double *y, int dp,
                                             PairMetrics &m) {
  const int n = DP > 0 ? DP : dp;
#pragma unroll
  for (int k = 0; k < n; ++k) {
    const double e = x[k] - y[k];
    const double e2 = e * e;
    m.g.r2 += e2;
    if (c.any_l1) m.g.l1 += fabs(e);
    if (c.any_ip) {
      m.g.dot = fma(x[k], y[k], m.g.dot);
      m.g.nx = fma(x[k], x[k], m.g.nx);
      m.g.ny = fma(y[k], y[k], m.g.ny);
    }
#pragma unroll
    for (int s = 0; s < MAX_ARD_SLOTS; ++s)
      if (s < c.nslots && k < MAX_ARD_DIM) m.ard[s] = fma(c.w[s][k], c.slot_l1[s] ? fabs(e) : e2, m.ard[s]);
  }
}

__device__ __forceinline__ double metric_of(const CompositeDev &c, const PairMetrics &m, int f) {
  const int id = c.metric[f];
  return id == 0 ? m.g.r2 : (id == 1 ? m.g.l1 : m.ard[id - 2]);
}

template <int FT>
__device__ __forceinline__ double composite_entry(const CompositeDev &c, const PairMetrics &m) {
  double v[FT];
  if (FactorLoop<FT>::unrolled) {
#pragma unroll
    for (int f = 0; f < FT; ++f) {
      if (f >= c.F) v[f] = 1.0;
      else if (kind_is_inner_product(c.cp[f].kind)) v[f] = cov_value_ip(c.cp[f], m.g);
      else v[f] = cov_value_shared(c.cp[f], metric_of(c, m, f));
    }
  } else {
    for (int f = 0; f < c.F; ++f) {
      if (kind_is_inner_product(c.cp[f].kind)) v[f] = cov_value_ip(c.cp[f], m.g);
      else v[f] = cov_value_rt(c.cp[f], metric_of(c, m, f));
    }
  }
  return composite_combine<FT>(c, v);
}

template <int DP, int FT>
__global__ void __launch_bounds__(ATHREADS) assemble_composite_kernel(CompositeDev c, double noise_abs,
                                                                      AssembleArgs a) {
  extern __shared__ __align__(16) double sm[];  // column points: AT x dp, row-major
  const int dp = DP > 0 ? DP : a.dp;
  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;

  int64_t tlin = blockIdx.x;
  int ti, tj;
  if (a.lower_only) {
    ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
    while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
    tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  } else {
    int mt = (int)(a.rows_p / AT);
    tj = (int)(tlin / mt);
    ti = (int)(tlin % mt);
  }
  int64_t m0 = (int64_t)ti * AT + a.row_off, n0 = (int64_t)tj * AT + a.col_off;
  const int64_t lm0 = m0, ln0 = n0;
  if (a.bc_T > 0) {
    m0 = ((int64_t)(lm0 / a.bc_T + a.bc_row0) * a.bc_Pr + a.bc_pr) * a.bc_T + lm0 % a.bc_T;
    n0 = ((int64_t)(ln0 / a.bc_T + a.bc_col0) * a.bc_Pc + a.bc_pc) * a.bc_T + ln0 % a.bc_T;
    if (n0 > m0 + AT - 1) return;
  }
  for (int i = tid; i < AT * dp; i += ATHREADS) sm[i] = a.X2[n0 * dp + i];

  const int64_t r0 = m0 + 2 * tx;
  const bool sym = a.symmetric != 0;
  double *out = a.out + (lm0 - a.row_off + 2 * tx) + (ln0 - a.col_off) * a.ld;
  const double *p0 = a.X1 + r0 * dp, *p1 = p0 + dp;
  double x0[DP > 0 ? DP : 1], x1[DP > 0 ? DP : 1];
  if (DP > 0) {
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      x0[k] = p0[k];
      x1[k] = p1[k];
    }
  }
  __syncthreads();
  for (int cc = 0; cc < AT / 4; ++cc) {
    const int j = ty + 4 * cc;
    const double *y = sm + j * dp;
    PairMetrics q0, q1;
    pair_metrics<DP>(c, DP > 0 ? x0 : p0, y, dp, q0);
    pair_metrics<DP>(c, DP > 0 ? x1 : p1, y, dp, q1);
    double k0 = composite_entry<FT>(c, q0), k1 = composite_entry<FT>(c, q1);
    if (q0.g.r2 == 0.0) k0 += noise_abs;
    if (q1.g.r2 == 0.0) k1 += noise_abs;
    const int64_t col = n0 + j;
    if (r0 >= a.rows || col >= a.cols) k0 = (sym && r0 == col) ? 1.0 : 0.0;
    if (r0 + 1 >= a.rows || col >= a.cols) k1 = (sym && r0 + 1 == col) ? 1.0 : 0.0;
    *reinterpret_cast<double2 *>(out + (int64_t)j * a.ld) = make_double2(k0, k1);
  }
}

template <int DP, int FT>
void launch_composite_one(cudaStream_t st, const CompositeDev &c, double noise_abs, const AssembleArgs &a,
                          unsigned tiles) {
  size_t smem = (size_t)AT * a.dp * sizeof(double);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(assemble_composite_kernel<DP, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  assemble_composite_kernel<DP, FT><<<tiles, ATHREADS, smem, st>>>(c, noise_abs, a);
}

template <int FT>
void launch_composite(cudaStream_t st, const CompositeDev &c, double noise_abs, const AssembleArgs &a, unsigned tiles) {
  if (!FactorLoop<FT>::unrolled) {  // many base kernels: the runtime-loop variant measured faster than the unrolled-DP one
    launch_composite_one<0, FT>(st, c, noise_abs, a, tiles);
    return;
  }
  switch (a.dp) {
    case 4: launch_composite_one<4, FT>(st, c, noise_abs, a, tiles); break;
    case 8: launch_composite_one<8, FT>(st, c, noise_abs, a, tiles); break;
    case 12: launch_composite_one<12, FT>(st, c, noise_abs, a, tiles); break;
    case 16: launch_composite_one<16, FT>(st, c, noise_abs, a, tiles); break;
    case 20: launch_composite_one<20, FT>(st, c, noise_abs, a, tiles); break;
    case 24: launch_composite_one<24, FT>(st, c, noise_abs, a, tiles); break;
    case 32: launch_composite_one<32, FT>(st, c, noise_abs, a, tiles); break;
    default: launch_composite_one<0, FT>(st, c, noise_abs, a, tiles); break;
  }
}

// ---------------------------------------------------------------------------------------------
// v2: Gram formulation on the FP64 tensor path with a lean epilogue.
//
// |x-y|^2 = |x|^2 + |y|^2 - 2 x.y comes out of DMMA directly by augmenting the contraction:
//   column point j (MMA row)   : [ -2 y_1 .. -2 y_d , 1     , |y|^2 , 0 , 0 ]
//   row point i    (MMA column): [    x_1 ..    x_d , |x|^2 , 1     , 0 , 0 ]
// so one 8x8x4 DMMA per 4 features replaces 64 subtract + 64 FMA instructions.  ncu of v1
// (profiles/ncu_assemble_v1_r1.txt) shows the kernel issue-bound (~176 thread instructions per
// Matern entry, FP64 pipe 43 % busy), so the epilogue is also rebuilt from FP64 FMAs only:
// exp() by Cody-Waite reduction + degree-13 Horner + exponent insertion, sqrt() from the FP32
// reciprocal-square-root seed + two Goldschmidt steps + one correction.  The MMA M index is the
// matrix column and the N index the matrix row, so each thread's accumulator pair is two adjacent
// rows: 16-byte stores.
// Cancellation: the Gram distance carries an absolute error ~ KA*eps*(|x|^2+|y|^2); pairs closer
// than 1e-4 of that scale (the diagonal, duplicates, near-duplicates) are recomputed from exact
// differences, which also keeps DiracKernel's |x-y| == 0 test exact.  The host only selects this
// path when the resulting relative error bound on a kernel entry is below 2e-13 (assemble()).
// ---------------------------------------------------------------------------------------------
// The fast epilogue works on V independent entries at a time, written operation by operation over
// the V lanes so that the compiler interleaves V dependency chains (the scalar version was bound by
// FP64 latency: ncu "wait" stalls, 46 % FP64-unit utilisation).
//
// exp(x), x <= 0: Cody-Waite reduction x = n ln2 + r, |r| <= 0.347, degree-13 Taylor polynomial in
// Estrin form (depth 4 instead of 13), exponent insertion.  Results below e^-700 ~ 1e-304 are flushed
// to zero (the library would return subnormals there).
template <int V>
__device__ __forceinline__ void fast_exp_nonpos_v(const double (&x)[V], double (&out)[V]) {
  const double L2E = 1.4426950408889634074, SHIFT = 6755399441055744.0;
  const double LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
  double r[V], r2[V], r4[V], r8[V], sc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double xc = fmax(x[v], -700.0);
    const double t = fma(xc, L2E, SHIFT);
    const int n = __double2loint(t);
    const double nd = t - SHIFT;
    r[v] = fma(nd, -LN2LO, fma(nd, -LN2HI, xc));
    sc[v] = __hiloint2double((n + 1023) << 20, 0);
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    r2[v] = r[v] * r[v];
    r4[v] = r2[v] * r2[v];
    r8[v] = r4[v] * r4[v];
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double a0 = fma(1.0, r[v], 1.0);                                          // 1/0! + r/1!
    const double a1 = fma(1.6666666666666666e-01, r[v], 0.5);                       // 1/2! + r/3!
    const double a2 = fma(8.333333333333333e-03, r[v], 4.1666666666666664e-02);     // 1/4! + r/5!
    const double a3 = fma(1.984126984126984e-04, r[v], 1.3888888888888889e-03);     // 1/6! + r/7!
    const double a4 = fma(2.7557319223985893e-06, r[v], 2.48015873015873e-05);      // 1/8! + r/9!
    const double a5 = fma(2.505210838544172e-08, r[v], 2.755731922398589e-07);      // 1/10! + r/11!
    const double a6 = fma(1.6059043836821613e-10, r[v], 2.08767569878681e-09);      // 1/12! + r/13!
    const double b0 = fma(a1, r2[v], a0), b1 = fma(a3, r2[v], a2), b2 = fma(a5, r2[v], a4);
    const double d0 = fma(b1, r4[v], b0), d1 = fma(a6, r4[v], b2);
    const double p = fma(d1, r8[v], d0);
    out[v] = (x[v] < -700.0) ? 0.0 : p * sc[v];
  }
}

__device__ __noinline__ double cold_sqrt(double a) { return sqrt(a); }

// sqrt(a) for moderate positive a: FP32 reciprocal-square-root seed, two Goldschmidt steps, one
// correction (a handful of FMAs instead of the library's special-case handling).
template <int V>
__device__ __forceinline__ void fast_sqrt_pos_v(const double (&a)[V], double (&out)[V]) {
  double g[V], h[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double y = (double)rsqrtf((float)a[v]);
    g[v] = a[v] * y;
    h[v] = 0.5 * y;
  }
#pragma unroll
  for (int it = 0; it < 2; ++it)
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const double e = fma(-g[v], h[v], 0.5);
      g[v] = fma(g[v], e, g[v]);
      h[v] = fma(h[v], e, h[v]);
    }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double d = fma(-g[v], g[v], a[v]);
    out[v] = fma(d, h[v], g[v]);
    if (!(a[v] > 1e-30 && a[v] < 1e30)) out[v] = cold_sqrt(a[v]);  // outside the FP32 seed's range
  }
}

// Hot-path covariance from scalars held in registers (no access to the by-value CovParams, which
// the compiler would otherwise spill to local memory).  Matern: orders p <= 3, coefficients
// right-aligned in m0..m3 (leading ones zero for smaller p).
struct FastCov {
  double a2, c1, c2, m0, m1, m2, m3;
};

template <int KIND, int V>
__device__ __forceinline__ void cov_value_fast_v(const FastCov &f, const double (&r2)[V], double (&k)[V]) {
  double arg[V], ex[V];
  if (KIND == DGP_KERNEL_MATERN) {
    double r[V], sum[V];
    fast_sqrt_pos_v<V>(r2, r);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const double u = f.c2 * r[v];
      double s = fma(f.m0, u, f.m1);
      s = fma(s, u, f.m2);
      sum[v] = fma(s, u, f.m3);
      arg[v] = -f.c1 * r[v];
    }
    fast_exp_nonpos_v<V>(arg, ex);
#pragma unroll
    for (int v = 0; v < V; ++v) k[v] = ex[v] * sum[v];
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) arg[v] = -r2[v] * f.c1;
    fast_exp_nonpos_v<V>(arg, ex);
#pragma unroll
    for (int v = 0; v < V; ++v) k[v] = f.a2 * ex[v];
  }
}

// Cold path: exact differences from the shared-memory operands (the column operand stores -2y).
template <int KIND>
__device__ __noinline__ double exact_entry(const CovParams *cp, const double *xr, const double *yc, int dp) {
  double d = 0.0;
  for (int c = 0; c < dp; ++c) {
    const double df = fma(0.5, yc[c], xr[c]);
    d = fma(df, df, d);
  }
  double k = cov_value<KIND>(*cp, d);
  if (d == 0.0) k += cp->noise_abs;
  return k;
}

__host__ __device__ constexpr int gram_ldk(int dp) { return (dp + 4 <= 20) ? 20 : 36; }  // row stride = 4 mod 16 doubles

template <int KIND, int DP>
__global__ void __launch_bounds__(ATHREADS, 2) assemble_gram_kernel(CovParams cp, AssembleArgs a) {
  constexpr int KA = DP + 4, LDK = gram_ldk(DP);
  extern __shared__ __align__(16) double sm[];
  double *Brow = sm;                 // row points i (MMA N):    [i][k], x | |x|^2 | 1 | 0 0
  double *Acol = sm + AT * LDK;      // column points j (MMA M): [j][k], -2y | 1 | |y|^2 | 0 0
  __shared__ double red[ATHREADS / 32];
  __shared__ double s_thr;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;  // per 64-column half: 2 column blocks of 32 x 4 row blocks of 32

  int64_t tlin = blockIdx.x;
  int ti, tj;
  if (a.diag_only && a.lower_only && a.bc_T == 0) {
    ti = (int)(a.rows_p / AT) - 1;  // the launch covers the last tile row (the padded one)
    tj = (int)tlin;
  } else if (a.lower_only) {
    ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
    while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
    tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  } else {
    int mt = (int)(a.rows_p / AT);
    tj = (int)(tlin / mt);
    ti = (int)(tlin % mt);
  }
  int64_t m0 = (int64_t)ti * AT + a.row_off, n0 = (int64_t)tj * AT + a.col_off;
  const int64_t lm0 = m0, ln0 = n0;
  if (a.bc_T > 0) {
    m0 = ((int64_t)(lm0 / a.bc_T + a.bc_row0) * a.bc_Pr + a.bc_pr) * a.bc_T + lm0 % a.bc_T;
    n0 = ((int64_t)(ln0 / a.bc_T + a.bc_col0) * a.bc_Pc + a.bc_pc) * a.bc_T + ln0 % a.bc_T;
    if (n0 > m0 + AT - 1) return;
  }
  // diag_only (edge pass): just the tiles assemble_gram_lean_kernel leaves out, those touching the padding
  if (a.diag_only && !(m0 + AT > a.rows || n0 + AT > a.cols)) return;

  // ---- augmented operands: thread p < 128 builds row point p, thread p >= 128 column point p-128 ----
  double nrm;
  {
    const bool is_row = tid < AT;
    const int pidx = is_row ? tid : tid - AT;
    const double *src = is_row ? (a.X1 + (m0 + pidx) * DP) : (a.X2 + (n0 + pidx) * DP);
    double *dst = (is_row ? Brow : Acol) + pidx * LDK;
    double v[DP];
    nrm = 0.0;
#pragma unroll
    for (int c = 0; c < DP / 2; ++c) {
      const double2 q = reinterpret_cast<const double2 *>(src)[c];
      v[2 * c] = q.x;
      v[2 * c + 1] = q.y;
    }
#pragma unroll
    for (int c = 0; c < DP; ++c) nrm = fma(v[c], v[c], nrm);
#pragma unroll
    for (int c = 0; c < DP; ++c) dst[c] = is_row ? v[c] : -2.0 * v[c];
    dst[DP] = is_row ? nrm : 1.0;
    dst[DP + 1] = is_row ? 1.0 : nrm;
    dst[DP + 2] = 0.0;
    dst[DP + 3] = 0.0;
  }
  // fix-up threshold from the largest squared norm in the tile
  {
    double mx = nrm;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
      double m2 = red[0];
      for (int w = 1; w < ATHREADS / 32; ++w) m2 = fmax(m2, red[w]);
      s_thr = fmax(2.0e-4 * m2, 1e-300);  // 1e-4 * (|x|^2 + |y|^2) at most; never below an exact zero
    }
    __syncthreads();
  }
  const double thr = s_thr;
  FastCov fc;
  fc.a2 = cp.a2;  fc.c1 = cp.c1;  fc.c2 = cp.c2;
  {
    const int p = cp.p;  // launch side guarantees p <= 3 for this kernel
    fc.m0 = (p == 3) ? cp.coef[0] : 0.0;
    fc.m1 = (p == 3) ? cp.coef[1] : (p == 2 ? cp.coef[0] : 0.0);
    fc.m2 = (p == 3) ? cp.coef[2] : (p == 2 ? cp.coef[1] : (p == 1 ? cp.coef[0] : 0.0));
    fc.m3 = (p == 3) ? cp.coef[3] : (p == 2 ? cp.coef[2] : (p == 1 ? cp.coef[1] : cp.coef[0]));
  }

  // ---- two column halves of 64; per half every warp owns 32 columns x 32 rows = 4 x 4 accumulator pairs ----
  const bool sym = a.symmetric != 0;
  const bool edge = (m0 + AT > a.rows) || (n0 + AT > a.cols);
  double *out = a.out + (lm0 - a.row_off) + (ln0 - a.col_off) * a.ld;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    const int jbase = half * 64 + wm * 32, ibase = wn * 32;
    const double *ap = Acol + (jbase + g) * LDK + t;
    const double *bp = Brow + (ibase + g) * LDK + t;
#pragma unroll
    for (int kk = 0; kk < KA / 4; ++kk) {
      double af[4], bf[4];
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) af[mb] = ap[mb * 8 * LDK + kk * 4];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) bf[nb] = bp[nb * 8 * LDK + kk * 4];
#pragma unroll
      for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(acc[mb][nb][0]), "+d"(acc[mb][nb][1])
                       : "d"(af[mb]), "d"(bf[nb]));
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const int jl = jbase + mb * 8 + g;
      const int64_t col = n0 + jl;
#pragma unroll
      for (int q = 0; q < 2; ++q) {  // two row blocks = four entries at a time
        double r2[4], kv[4];
        bool exact[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double v = acc[mb][2 * q + (u >> 1)][u & 1];
          exact[u] = !(v > thr);  // cancellation-prone (diagonal, duplicates, near-duplicates)
          r2[u] = fmax(v, thr);
        }
        cov_value_fast_v<KIND, 4>(fc, r2, kv);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int il = ibase + (2 * q + (u >> 1)) * 8 + 2 * t + (u & 1);
          if (exact[u]) kv[u] = exact_entry<KIND>(&cp, Brow + il * LDK, Acol + jl * LDK, DP);
          if (edge) {
            const int64_t row = m0 + il;
            if (row >= a.rows || col >= a.cols) kv[u] = (sym && row == col) ? 1.0 : 0.0;
          }
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int il = ibase + (2 * q + h2) * 8 + 2 * t;
          *reinterpret_cast<double2 *>(out + il + (int64_t)jl * a.ld) = make_double2(kv[2 * h2], kv[2 * h2 + 1]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// v3: lean Gram epilogue for tiles in which no pair can be at distance exactly zero (every tile of
// a cross matrix without noise; the strictly-lower tiles of a training matrix whose points are
// pairwise distinct -- Data::has_dup is established once on the host).  There the only use of the
// exact-difference fallback, DiracKernel's |x-y| == 0 test, cannot fire, and the relative error of
// an entry is bounded by the same slope * KA * eps * 2 max|x|^2 the host gate checks (k is smooth
// in r^2, so a small r^2 does not amplify it).  Diagonal tiles go to assemble_gram_kernel.
//
// Per entry: 1 DADD (|x|^2 + |y|^2 seeds the accumulator) + DP DMMA slots (-2 x.y) + epilogue of
//   SE      x = -c r^2; n = rint(64 x / ln 2); r = x - n ln2/64 (two FMAs); degree-5 polynomial;
//           a^2 2^(j/64) from a 64-entry shared-memory table; exponent added as an integer: 10 DFMA/DMUL/DADD
//   Matern  clamp; MUFU.RSQ64H seed + one Goldschmidt step + one Newton correction (7); polynomial in r (3);
//           the exponential as above (10) and one product
// against ~45 (SE) / ~75 (Matern) instructions in v2's epilogue.  16 columns x 32 rows per warp and pass
// keep the accumulators at 32 registers so that three CTAs fit per SM.
// ---------------------------------------------------------------------------------------------
// Staged (bulk / TMA) store variant, option "assembly_store" = 1: each 32-column chunk of the tile is written to a
// shared-memory staging buffer (column stride 136 doubles: conflict-free 16-byte stores, 16-byte aligned columns) and
// one elected thread hands the 32 columns to the bulk-copy engine as 1 KB cp.async.bulk.global.shared::cta stores.
constexpr int STG_LD = AT + 8;                       // doubles per staged column
constexpr int STG_DOUBLES = 32 * STG_LD;             // one chunk

// P3: Matern order 3 (cubic polynomial in r); orders 1 and 2 share the quadratic form.
template <int KIND, int DP, bool P3, bool BULK = false>
__global__ void __launch_bounds__(ATHREADS, 3)
    assemble_gram_lean_kernel(LeanCov f, LeanConst lc, ExpTab T, double kdiag, AssembleArgs a) {
  constexpr int LDK = lean_ldk<DP>();
  extern __shared__ __align__(16) double sm_raw[];
  double *tab = sm_raw;                 // 64 entries x 16 copies
  double *Brow = tab + 64 * 16;         // row points i (MMA N):    [i][k] = x
  double *Acol = Brow + AT * LDK;       // column points j (MMA M): [j][k] = -2 y
  double *nx = Acol + AT * LDK;         // |x_i|^2
  double *ny = nx + AT;                 // |y_j|^2
  double *stg = ny + AT;                // BULK: staging buffer of one 32-column chunk

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;

  int64_t tlin = blockIdx.x;
  int ti, tj;
  if (a.lower_only) {
    // FP32 estimate of the triangular root, corrected by the two loops (no FP64 sqrt chain in the prologue)
    ti = (int)((sqrtf(8.0f * (float)tlin + 1.0f) - 1.0f) * 0.5f);
    while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
    tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  } else {
    int mt = (int)(a.rows_p / AT);
    tj = (int)(tlin / mt);
    ti = (int)(tlin % mt);
  }
  int64_t m0 = (int64_t)ti * AT + a.row_off, n0 = (int64_t)tj * AT + a.col_off;
  const int64_t lm0 = m0, ln0 = n0;
  if (a.bc_T > 0) {
    m0 = ((int64_t)(lm0 / a.bc_T + a.bc_row0) * a.bc_Pr + a.bc_pr) * a.bc_T + lm0 % a.bc_T;
    n0 = ((int64_t)(ln0 / a.bc_T + a.bc_col0) * a.bc_Pc + a.bc_pc) * a.bc_T + ln0 % a.bc_T;
    if (n0 > m0 + AT - 1) return;
  }
  // tiles touching the padding are left to assemble_gram_kernel (edge pass)
  if (m0 + AT > a.rows || n0 + AT > a.cols) return;
  // diagonal tile of a symmetric matrix: the launch side guarantees that only i == j is at distance zero
  // (or that there is no noise term), so those entries are simply k(0) + noise
  const bool diag_tile = a.symmetric && m0 == n0;

  {
    const bool is_row = tid < AT;
    const int pidx = is_row ? tid : tid - AT;
    const double *src = is_row ? (a.X1 + (m0 + pidx) * DP) : (a.X2 + (n0 + pidx) * DP);
    double *dst = (is_row ? Brow : Acol) + pidx * LDK;
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
    const double sc = is_row ? 1.0 : -2.0;
#pragma unroll
    for (int c = 0; c < DP / 2; ++c)
      reinterpret_cast<double2 *>(dst)[c] = make_double2(sc * v[2 * c], sc * v[2 * c + 1]);
    (is_row ? nx : ny)[pidx] = nrm;
#pragma unroll
    for (int e = 0; e < 4; ++e) tab[e * 256 + tid] = T.v[(e * 256 + tid) >> 4];   // consecutive lanes, consecutive words
  }
  __syncthreads();

  double *out = a.out + (lm0 - a.row_off) + (ln0 - a.col_off) * a.ld;
  const int ibase = wn * 32;
  const double *bp = Brow + (ibase + g) * LDK + t;
  const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab) + ((unsigned)(lane & 15) << 3);  // this lane's copy
#pragma unroll 1
  for (int chunk = 0; chunk < 4; ++chunk) {
    if (BULK && chunk > 0) {  // the engine must have read the previous chunk out of the staging buffer
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      __syncthreads();
    }
    const int jbase = chunk * 32 + wm * 16;
    const double *ap = Acol + (jbase + g) * LDK + t;
    double acc[2][4][2];
    {
      const double y0 = ny[jbase + g], y1 = ny[jbase + 8 + g];
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const double2 xn = *reinterpret_cast<const double2 *>(nx + ibase + nb * 8 + 2 * t);
        acc[0][nb][0] = y0 + xn.x;  acc[0][nb][1] = y0 + xn.y;
        acc[1][nb][0] = y1 + xn.x;  acc[1][nb][1] = y1 + xn.y;
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
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const int jl = jbase + mb * 8 + g;
      double *ocol = BULK ? (stg + (jl - chunk * 32) * STG_LD + ibase + 2 * t) : (out + (int64_t)jl * a.ld + ibase + 2 * t);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double x[4], kv[4];
        if (KIND == DGP_KERNEL_MATERN) {
          double r[4], s[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // clamp to >= ~1e-290 on the high word (an integer max: negative and tiny values alike)
            const double av = acc[mb][2 * q + (u >> 1)][u & 1];
            const double v = __hiloint2double(max(__double2hiint(av), 0x03c00000), __double2loint(av));
            // sqrt: 2^-22 seed, one Goldschmidt step (-> 2^-43), one Newton correction with the seed's
            // half reciprocal (its 2^-22 error only scales the 2^-43 residual)
            const double y = rsqrt_seed(v);
            double gg = v * y;
            const double hh = 0.5 * y;
            const double e = fma(-gg, hh, 0.5);
            gg = fma(gg, e, gg);
            const double dd = fma(-gg, gg, v);
            r[u] = fma(dd, hh, gg);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            double p = P3 ? fma(f.m0, r[u], f.m1) : f.m1;
            p = fma(p, r[u], f.m2);
            s[u] = fma(p, r[u], f.m3);
            x[u] = -r[u];
          }
          lean_exp_v<4>(x, lc, tab_s, kv);
#pragma unroll
          for (int u = 0; u < 4; ++u) kv[u] *= s[u];
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = -acc[mb][2 * q + (u >> 1)][u & 1];
          lean_exp_v<4>(x, lc, tab_s, kv);
        }
        *reinterpret_cast<double2 *>(ocol + (2 * q) * 8) = make_double2(kv[0], kv[1]);
        *reinterpret_cast<double2 *>(ocol + (2 * q + 1) * 8) = make_double2(kv[2], kv[3]);
      }
    }
    if (BULK) {
      // the diagonal of a symmetric tile is k(0) + noise: patched in the staging buffer before it leaves
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // staged generic writes -> async-proxy reads
      __syncthreads();
      if (diag_tile && tid < 32) stg[tid * STG_LD + chunk * 32 + tid] = kdiag;
      if (diag_tile) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
      }
      if (tid == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(stg);
        double *dst = out + (int64_t)(chunk * 32) * a.ld;
#pragma unroll 8
        for (int c = 0; c < 32; ++c)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst + (int64_t)c * a.ld),
                       "r"(src + (unsigned)(c * STG_LD * 8)), "n"(AT * 8)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    }
  }
  if (BULK) {
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // shared memory is released at exit
    return;
  }
  if (diag_tile) {  // CTA-uniform; the barrier orders these stores after the tile's own
    __syncthreads();
    if (tid < AT) out[tid + (int64_t)tid * a.ld] = kdiag;
  }
}

template <int KIND, int DP>
void launch_lean(cudaStream_t st, const CovParams &cp, const AssembleArgs &a, unsigned tiles, bool bulk) {
  LeanCov f;
  f.scale = (KIND == DGP_KERNEL_MATERN) ? cp.c1 : sqrt(cp.c1);
  f.m0 = f.m1 = f.m2 = f.m3 = 0.0;
  if (KIND == DGP_KERNEL_MATERN) {
    // sum_i coef[i] (c2 r)^(p-i) in terms of u = c1 r (c2 = 2 c1): coef[i] 2^(p-i) u^(p-i), right-aligned in m0..m3
    double *m[4] = {&f.m0, &f.m1, &f.m2, &f.m3};
    for (int i = 0; i <= cp.p; ++i) *m[3 - (cp.p - i)] = cp.coef[i] * ldexp(1.0, cp.p - i);
  }
  const double a2 = (KIND == DGP_KERNEL_MATERN) ? 1.0 : cp.a2;
  const ExpTab T = lean_exp_table(a2);
  const LeanConst lc = lean_constants();
  // value at distance zero plus the noise term (DiracKernel.scala:43-47 adds |sigma| where |x-y| == 0)
  const double kdiag = ((KIND == DGP_KERNEL_MATERN) ? cp.coef[cp.p] : cp.a2) + cp.noise_abs;
  // bulk stores need 16-byte aligned destination columns (ld even is already required; out from cudaMalloc)
  bulk = bulk && ((uintptr_t)a.out % 16 == 0) && (a.ld % 2 == 0);
  size_t smem = ((size_t)2 * AT * lean_ldk<DP>() + 2 * AT + 64 * 16 + (bulk ? STG_DOUBLES : 0)) * sizeof(double);
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<tiles, ATHREADS, smem, st>>>(f, lc, T, kdiag, a);
  };
  if constexpr (KIND == DGP_KERNEL_MATERN) {
    if (cp.p == 3) {
      go(assemble_gram_lean_kernel<KIND, DP, true>);
      return;
    }
  }
  if constexpr (KIND == DGP_KERNEL_SE && (DP == 8 || DP == 16)) {
    if (bulk) {
      go(assemble_gram_lean_kernel<KIND, DP, false, true>);
      return;
    }
  }
  go(assemble_gram_lean_kernel<KIND, DP, false>);
}

template <int KIND>
bool launch_lean_kind(cudaStream_t st, const CovParams &cp, const AssembleArgs &a, unsigned tiles, bool bulk = false) {
  switch (a.dp) {
    case 4: launch_lean<KIND, 4>(st, cp, a, tiles, bulk); return true;
    case 8: launch_lean<KIND, 8>(st, cp, a, tiles, bulk); return true;
    case 12: launch_lean<KIND, 12>(st, cp, a, tiles, bulk); return true;
    case 16: launch_lean<KIND, 16>(st, cp, a, tiles, bulk); return true;
    case 20: launch_lean<KIND, 20>(st, cp, a, tiles, bulk); return true;
    case 24: launch_lean<KIND, 24>(st, cp, a, tiles, bulk); return true;
    case 32: launch_lean<KIND, 32>(st, cp, a, tiles, bulk); return true;
    default: return false;
  }
}

template <int KIND, int DP>
void launch_gram(cudaStream_t st, const CovParams &cp, const AssembleArgs &a, unsigned tiles) {
  size_t smem = (size_t)2 * AT * gram_ldk(DP) * sizeof(double);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(assemble_gram_kernel<KIND, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  assemble_gram_kernel<KIND, DP><<<tiles, ATHREADS, smem, st>>>(cp, a);
}

template <int KIND>
bool launch_gram_kind(cudaStream_t st, const CovParams &cp, const AssembleArgs &a, unsigned tiles) {
  switch (a.dp) {
    case 4: launch_gram<KIND, 4>(st, cp, a, tiles); return true;
    case 8: launch_gram<KIND, 8>(st, cp, a, tiles); return true;
    case 12: launch_gram<KIND, 12>(st, cp, a, tiles); return true;
    case 16: launch_gram<KIND, 16>(st, cp, a, tiles); return true;
    case 20: launch_gram<KIND, 20>(st, cp, a, tiles); return true;
    case 24: launch_gram<KIND, 24>(st, cp, a, tiles); return true;
    case 32: launch_gram<KIND, 32>(st, cp, a, tiles); return true;
    default: return false;
  }
}

template <int KIND, int DP>
void launch_one(cudaStream_t st, const CovParams &cp, const AssembleArgs &a, unsigned tiles) {
  size_t smem = (size_t)AT * a.dp * sizeof(double);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(assemble_kernel<KIND, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  assemble_kernel<KIND, DP><<<tiles, ATHREADS, smem, st>>>(cp, a);
}

template <int KIND>
void launch_kind(cudaStream_t st, const CovParams &cp, const AssembleArgs &a, unsigned tiles) {
  switch (a.dp) {
    case 4: launch_one<KIND, 4>(st, cp, a, tiles); break;
    case 8: launch_one<KIND, 8>(st, cp, a, tiles); break;
    case 12: launch_one<KIND, 12>(st, cp, a, tiles); break;
    case 16: launch_one<KIND, 16>(st, cp, a, tiles); break;
    case 20: launch_one<KIND, 20>(st, cp, a, tiles); break;
    case 24: launch_one<KIND, 24>(st, cp, a, tiles); break;
    case 32: launch_one<KIND, 32>(st, cp, a, tiles); break;
    default: launch_one<KIND, 0>(st, cp, a, tiles); break;
  }
}

}  // namespace

int32_t assemble(Ctx *ctx, cudaStream_t st, const CovParams &cp, const AssembleArgs &a) {
  if (a.rows_p % AT || a.cols_p % AT || a.dp % 4 || a.dp <= 0 || (a.ld & 1))
    return ctx->fail(DGP_ERR_SHAPE, "assemble: padded sizes must be multiples of 128, dp of 4, ld even");
  if ((size_t)AT * a.dp * sizeof(double) > 200 * 1024)
    return ctx->fail(DGP_ERR_SHAPE, "assemble: at most 200 features are supported");
  int64_t mt = a.rows_p / AT, nt = a.cols_p / AT;
  if (a.lower_only && mt != nt) return ctx->fail(DGP_ERR_SHAPE, "assemble: lower_only needs a square matrix");
  int64_t tiles = a.lower_only ? mt * (mt + 1) / 2 : mt * nt;
  if (tiles == 0) return DGP_OK;
  if (cp.kind == DGP_KERNEL_COMPOSITE) {
    if (!cp.comp) return ctx->fail(DGP_ERR_ARG, "assemble: composite kernel without an expression");
    if (cp.comp->F <= 4) launch_composite<4>(st, *cp.comp, cp.noise_abs, a, (unsigned)tiles);
    else launch_composite<MAXF>(st, *cp.comp, cp.noise_abs, a, (unsigned)tiles);
    ctx->launches++;
    DGP_CUDA_OK(ctx, cudaGetLastError());
    return DGP_OK;
  }
  bool done = false;
  const bool se_like = cp.kind == DGP_KERNEL_SE || cp.kind == DGP_KERNEL_RBF || cp.kind == DGP_KERNEL_ARD_SE;
  const bool matern_ok = cp.kind == DGP_KERNEL_MATERN && cp.p <= 3;
  if (a.use_gram && (se_like || matern_ok)) {
    // lean epilogue wherever no pair can be at distance zero; the amplitude is folded into its exp table,
    // so it must be far from the exponent limits
    const bool lean = ctx->assembly != 3 && (cp.noise_abs == 0.0 || (a.symmetric && a.nodup)) &&
                      cp.a2 >= 1e-150 && cp.a2 <= 1e150;
    if (lean) {
      done = se_like ? launch_lean_kind<DGP_KERNEL_SE>(st, cp, a, (unsigned)tiles, ctx->assembly_store == 1)
                     : launch_lean_kind<DGP_KERNEL_MATERN>(st, cp, a, (unsigned)tiles);
      const bool padded = a.rows_p != a.rows || a.cols_p != a.cols;
      if (done && padded) {  // edge pass: tiles touching the padding, with the checked epilogue
        AssembleArgs ad = a;
        ad.diag_only = 1;
        const unsigned etiles = (a.lower_only && a.bc_T == 0) ? (unsigned)mt : (unsigned)tiles;
        if (se_like) launch_gram_kind<DGP_KERNEL_SE>(st, cp, ad, etiles);
        else launch_gram_kind<DGP_KERNEL_MATERN>(st, cp, ad, etiles);
        ctx->launches++;
      }
    } else if (a.dp >= 16 || ctx->assembly >= 2) {
      // measured on B200 (profiles/assemble_r1.json): with the checked epilogue the Gram kernel only wins
      // from 16 features up (Matern-5/2 d = 16: 0.88 vs 1.08 ms at N = 16384; SE d = 8: 0.57 vs 0.42 ms)
      done = se_like ? launch_gram_kind<DGP_KERNEL_SE>(st, cp, a, (unsigned)tiles)
                     : launch_gram_kind<DGP_KERNEL_MATERN>(st, cp, a, (unsigned)tiles);
    }
  }
  if (!done) switch (cp.kind) {
    case DGP_KERNEL_SE:
    case DGP_KERNEL_RBF:
    case DGP_KERNEL_ARD_SE: launch_kind<DGP_KERNEL_SE>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_MATERN: launch_kind<DGP_KERNEL_MATERN>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_PERIODIC: launch_kind<DGP_KERNEL_PERIODIC>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_DIRAC: launch_kind<DGP_KERNEL_DIRAC>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_CAUCHY: launch_kind<DGP_KERNEL_CAUCHY>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_LAPLACIAN: launch_kind<DGP_KERNEL_LAPLACIAN>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_RATQUAD: launch_kind<DGP_KERNEL_RATQUAD>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_TSTUDENT: launch_kind<DGP_KERNEL_TSTUDENT>(st, cp, a, (unsigned)tiles); break;
    default: return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "assemble: kernel kind %d", cp.kind);
  }
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

}  // namespace dgp

namespace dgp {

bool gram_is_accurate(const CovParams &cp, int dp, double max_sq_norm) {
  if (dp > 32 || !(max_sq_norm >= 0.0) || !std::isfinite(max_sq_norm)) return false;
  double slope;  // max |d log k / d r^2|
  switch (cp.kind) {
    case DGP_KERNEL_SE:
    case DGP_KERNEL_RBF:
    case DGP_KERNEL_ARD_SE: slope = cp.c1; break;
    case DGP_KERNEL_MATERN:
      if (cp.p < 1) return false;        // exp(-r/l) is not smooth in r^2 at 0
      slope = 0.5 * cp.c1 * cp.c1;       // <= nu / l^2, attained at r -> 0 for p = 1
      break;
    default: return false;
  }
  if (!std::isfinite(slope)) return false;
  const double bound = slope * (dp + 4) * 1.2e-16 * 2.0 * max_sq_norm;
  return bound <= 2e-13;
}

}  // namespace dgp
