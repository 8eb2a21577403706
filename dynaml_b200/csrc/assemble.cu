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
#include "assemble.cuh"

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
    m0 = ((int64_t)(lm0 / a.bc_T) * a.bc_Pr + a.bc_pr) * a.bc_T + lm0 % a.bc_T;
    n0 = ((int64_t)(ln0 / a.bc_T) * a.bc_Pc + a.bc_pc) * a.bc_T + ln0 % a.bc_T;
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
        if (KIND == DGP_KERNEL_PERIODIC) {
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
        if (KIND == DGP_KERNEL_PERIODIC) {
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

__global__ void scale_points_kernel(const double *X, double *Xs, const double *inv_b, int64_t n, int dp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * dp) Xs[i] = X[i] * inv_b[i % dp];
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
  switch (cp.kind) {
    case DGP_KERNEL_SE:
    case DGP_KERNEL_RBF:
    case DGP_KERNEL_ARD_SE: launch_kind<DGP_KERNEL_SE>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_MATERN: launch_kind<DGP_KERNEL_MATERN>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_PERIODIC: launch_kind<DGP_KERNEL_PERIODIC>(st, cp, a, (unsigned)tiles); break;
    case DGP_KERNEL_DIRAC: launch_kind<DGP_KERNEL_DIRAC>(st, cp, a, (unsigned)tiles); break;
    default: return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "assemble: kernel kind %d", cp.kind);
  }
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t scale_points(Ctx *ctx, cudaStream_t st, const double *X, double *Xs, const double *d_inv_b, int64_t n,
                     int dp) {
  int64_t total = n * dp;
  if (total == 0) return DGP_OK;
  scale_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, Xs, d_inv_b, n, dp);
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

}  // namespace dgp
