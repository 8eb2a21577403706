// solve.cuh -- memory-bound vector kernels of the path: blocked triangular solves, reductions,
// GEMV, mirror / pack helpers.
#pragma once
#include "common.cuh"

namespace dgp {

int32_t solve_init(Ctx *ctx);  // opt in to the dynamic shared memory of the chained solves, once per context

// x <- L^-1 x (forward) using the diagonal-block inverses; n % 128 == 0.
// recLTriagSolve, CORE/algebra/PartitionedMatrixSolvers.scala:28-63.
int32_t trsv_lower_fwd(Ctx *ctx, cudaStream_t st, const double *L, int64_t ldl, const double *inv, int64_t n,
                       double *x);
// x <- L^-T x (backward).  recUTriagSolve, PartitionedMatrixSolvers.scala:66-109.
int32_t trsv_lower_bwd(Ctx *ctx, cudaStream_t st, const double *L, int64_t ldl, const double *inv, int64_t n,
                       double *x);
// out[0] = sum_i log A(i,i)   (btrace(blog(L)), CORE/algebra/package.scala:101-108,147-149)
int32_t diag_log_sum(Ctx *ctx, cudaStream_t st, const double *A, int64_t lda, int64_t n, double *out);
// out[0] = x . y   (PartitionedMatrixOps.scala:326-335)
int32_t dot(Ctx *ctx, cudaStream_t st, const double *x, const double *y, int64_t n, double *out);
// y[j] = beta_y[j] + sum_i A(i,j) x[i], j < ncols: y = base + A^T x  (crossKernel.t * alpha)
int32_t gemv_t(Ctx *ctx, cudaStream_t st, const double *A, int64_t lda, int64_t rows, int64_t cols, const double *x,
               const double *base /* nullable */, double *y);
// bar[i] = sum_{j <= i} L(i,j)   (root * ones, BlockedMultiVariateGaussian.scala:58-68)
int32_t lower_row_sums(Ctx *ctx, cudaStream_t st, const double *L, int64_t ldl, int64_t n, double *bar);
// copy the lower triangle onto the upper one, tile by tile (n % 32 == 0)
int32_t mirror_lower(Ctx *ctx, cudaStream_t st, double *A, int64_t lda, int64_t n);
// dst <- src for an n x m block (both column-major)
int32_t copy_block(Ctx *ctx, cudaStream_t st, const double *src, int64_t lds, double *dst, int64_t ldd, int64_t n,
                   int64_t m);

}  // namespace dgp
