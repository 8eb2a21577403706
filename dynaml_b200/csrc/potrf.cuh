// potrf.cuh -- blocked Cholesky, triangular inverse and L^-T L^-1 on device-resident matrices.
#pragma once
#include "common.cuh"

namespace dgp {

int32_t potrf_init(Ctx *ctx);

// Outer block of the blocked Cholesky for an n x n matrix: ctx->nb when set, else chosen from n.
int64_t potrf_outer_block(const Ctx *ctx, int64_t n);

// In-place right-looking blocked Cholesky of the lower triangle of A (n x n, n % 128 == 0).
//   A    : on exit L in the lower triangle (diagonal blocks have their upper part zeroed;
//          tiles strictly above the diagonal are never touched)
//   inv  : n/128 blocks of 128x128 (column-major, ld = 128): inverse of each diagonal block of L
//   info : device int; set to (first failing global pivot index + 1) when a pivot is <= 0 or NaN
//          (only the first failure is kept); must be zeroed by the caller.
// Replaces bcholesky / choleskyPAcc (CORE/algebra/bcholesky.scala:85-148).
int32_t potrf_blocked(Ctx *ctx, double *A, int64_t lda, int64_t n, double *inv, int *info);
// Same, on explicit streams: S carries the trailing updates, P the panels (P == S: no look-ahead).
int32_t potrf_blocked_on(Ctx *ctx, cudaStream_t S, cudaStream_t P, double *A, int64_t lda, int64_t n, double *inv,
                         int *info, const char *scratch_name);

// potrf_blocked_on through a CUDA graph.  The first call for a given (A, lda, n, inv, info, S, P) runs eagerly (it
// allocates the scratch panel and builds the GEMM tile tables), the second captures the same launch sequence on the
// two streams into a graph, and from then on one cudaGraphLaunch on S replaces the ~5 launches per 128 columns.
// Same kernels, same order, same results.  Falls back to the eager path when ctx->use_graphs == 0.
int32_t potrf_blocked_graph(Ctx *ctx, cudaStream_t S, cudaStream_t P, double *A, int64_t lda, int64_t n, double *inv,
                            int *info, const char *scratch_name);

// W (lower) = L^-1 from L and the diagonal-block inverses; scratch is an n x n buffer (ld = lds)
// whose lower triangle is clobbered.  Tiles of W strictly above the diagonal are not written.
int32_t trtri_lower(Ctx *ctx, const double *L, int64_t ldl, const double *inv, int64_t n, double *W, int64_t ldw,
                    double *scratch, int64_t lds);
int32_t trtri_lower_on(Ctx *ctx, cudaStream_t S, const double *L, int64_t ldl, const double *inv, int64_t n, double *W,
                       int64_t ldw, double *scratch, int64_t lds);

// V (m x n) = B L^-T, i.e. the transpose of L^-1 B^T, for B (m x n, m and n multiples of 128; overwritten with
// partially updated columns): blocked substitution on tensor cores using the diagonal-block inverses; rank-128
// updates inside blocks of `nb` columns, one rank-nb update of every column beyond.  Solving from the right keeps every
// product in the NT form of the bulk-copy GEMM (the left-sided form needs NN products, and V^T V afterwards a TN one).
// recLTriagMultiSolve, CORE/algebra/PartitionedMatrixSolvers.scala:120-169.
// With P != S the block's own 128-column steps run on P beside the update of the columns beyond the next block.
int32_t trsm_right_lower_t(Ctx *ctx, cudaStream_t S, cudaStream_t P, const double *L, int64_t ldl, const double *inv,
                           int64_t n, double *B, int64_t ldb, double *V, int64_t ldv, int64_t m, int64_t nb);

// Kinv (lower tiles) = W^T W.
// The same inverse stored transposed, U = L^-T (upper triangular), and K^-1 = U U^T from it: with U every product of the
// pair is NT or NN (MN-major A), so the whole inverse runs on the bulk-copy GEMM kernel; W^T W is a TN product.
int32_t trtri_lower_transposed_on(Ctx *ctx, cudaStream_t S, const double *L, int64_t ldl, const double *inv, int64_t n,
                                  double *U, int64_t ldu, double *scratch, int64_t lds);
int32_t lauum_upper_on(Ctx *ctx, cudaStream_t S, const double *U, int64_t ldu, int64_t n, double *Kinv, int64_t ldk);
int32_t lauum_lower(Ctx *ctx, const double *W, int64_t ldw, int64_t n, double *Kinv, int64_t ldk);
int32_t lauum_lower_on(Ctx *ctx, cudaStream_t S, const double *W, int64_t ldw, int64_t n, double *Kinv, int64_t ldk);

}  // namespace dgp
