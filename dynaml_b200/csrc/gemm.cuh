// gemm.cuh -- interface of the FP64 tensor-core (DMMA) GEMM that carries every O(N^3) step of the
// path: Cholesky panel and trailing updates, the triangular inverse, L^-T L^-1, and the
// prediction TRSM/SYRK.
#pragma once
#include "common.cuh"

namespace dgp {

// Per-tile restriction of the contraction range, used to skip the structurally-zero half of a
// triangular operand.  (m0, n0) is the origin of the 128x64 output tile.
enum KMode {
  K_FULL = 0,
  K_GE_COL = 1,      // k in [n0, K)        : op(B) is lower triangular  (B[k][n] = 0 for k < n)
  K_LT_ROWEND = 2,   // k in [0, m0 + 128)  : op(A) is lower triangular  (A[m][k] = 0 for k > m)
  K_GE_ROW = 3,      // k in [m0, K)        : op(A) is upper triangular  (A[m][k] = 0 for k < m)
  K_LT_COLEND = 4    // k in [0, n0 + 64)   : op(B) is upper triangular  (B[k][n] = 0 for k > n)
};

struct GemmArgs {
  // C[M x N] = alpha * op(A)[M x K] * op(B)[K x N] + beta * C, everything column-major.
  //   a_kmajor == 0: op(A)(m,k) = A[m + k*lda]   (A stored M x K)
  //   a_kmajor == 1: op(A)(m,k) = A[k + m*lda]   (A stored K x M, i.e. op(A) = A^T)
  //   b_kmajor == 0: op(B)(k,n) = B[n + k*ldb]   (B stored N x K, i.e. op(B) = B^T)
  //   b_kmajor == 1: op(B)(k,n) = B[k + n*ldb]   (B stored K x N)
  const double *A = nullptr;
  const double *B = nullptr;
  double *C = nullptr;
  int64_t lda = 0, ldb = 0, ldc = 0;
  int M = 0, N = 0, K = 0;  // multiples of 128, 64, 16
  double alpha = 1.0, beta = 0.0;
  int a_kmajor = 0, b_kmajor = 0;
  int lower_only = 0;  // square C: only tiles with m0 >= n0
  int kmode = K_FULL;
  int reverse = 0;     // launch heavy tiles first when the k-range grows with the tile index
  int batch = 1;
  int no_prefetch = 0; // testing: skip the L2 prefetch of the C tile (makes the epilogue slow; stresses stage hand-over)
  int zero = 0;        // always 0: a value the compiler cannot fold (see the `empty` arrival in gemm.cu)
  int handover = 0;    // set by gemm() from ctx->gemm_handover
  int64_t strideA = 0, strideB = 0, strideC = 0;
  // Optional 2D block-cyclic mask (dist.cu): C is this rank's local part of a distributed
  // lower-triangular update, stored as T x T tiles; local tile (li, lj) is global tile
  // (li*Pr + pr, lj*Pc + pc).  Output tiles entirely above the *global* diagonal return at once.
  int bc_T = 0, bc_Pr = 1, bc_Pc = 1, bc_pr = 0, bc_pc = 0;
  int bc_row0 = 0, bc_col0 = 0;  // local tile index (units of T) of the first row / column of C
};

// Enqueue on `stream`.  Returns DGP_OK or a CUDA error recorded on ctx.
int32_t gemm(Ctx *ctx, cudaStream_t stream, const GemmArgs &args);
// Build (and cache) the tile table an M x N product will use, so that a later gemm() of that shape neither
// allocates nor copies inside a stream pipeline.
int32_t gemm_prepare(Ctx *ctx, int M, int N, int lower_only, int reverse);
int32_t gemm_init(Ctx *ctx);  // opt in to the large dynamic shared memory once per context
void gemm_release(Ctx *ctx);  // free the cached tile tables of a context

}  // namespace dgp
