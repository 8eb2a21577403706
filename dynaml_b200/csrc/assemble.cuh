// assemble.cuh -- kernel-matrix assembly interface.
#pragma once
#include "common.cuh"
#include "covfn.cuh"

namespace dgp {

struct AssembleArgs {
  // out(i - row_off, j - col_off) = k(X1[i], X2[j]) for the tile grid covering
  // [row_off, row_off + rows_p) x [col_off, col_off + cols_p); point sets are row-major n x dp on
  // the device (dp a multiple of 4, zero padded) and long enough for the padded ranges.
  const double *X1 = nullptr;
  const double *X2 = nullptr;
  int dp = 0;
  int64_t rows = 0, cols = 0;       // real extents (global indices >= these are padding)
  int64_t rows_p = 0, cols_p = 0;   // extents of the produced block, multiples of 128
  int64_t row_off = 0, col_off = 0;
  double *out = nullptr;
  int64_t ld = 0;
  int symmetric = 0;   // X1 == X2 as point sets: padding becomes the identity, noise applies
  int lower_only = 0;  // only tiles on/below the diagonal (requires rows_p == cols_p, offsets equal)
  // Optional 2D block-cyclic layout (dist.cu): `out` is a rank's local rows_p x cols_p matrix of
  // T x T tiles; local tile (li, lj) holds global tile (li*Pr + pr, lj*Pc + pc).  Tiles entirely
  // above the global diagonal are skipped.
  int bc_T = 0, bc_Pr = 1, bc_Pc = 1, bc_pr = 0, bc_pc = 0;
  int bc_row0 = 0, bc_col0 = 0;  // local tile index (units of bc_T) of the first row / column of `out`
  // 1: Gram/DMMA distance kernel (SE, RBF, ARD-SE, Matern with dp <= 32); the caller sets it only when
  // gram_is_accurate() holds.  0: exact-difference kernel.
  int use_gram = 0;
  // 1: the points of a symmetric assembly are pairwise distinct (Data::has_dup == false and the ARD
  // scaling cannot collapse them), so zero distances occur on the diagonal only and the strictly-lower
  // tiles may take the lean Gram epilogue even with a Dirac noise term.
  int nodup = 0;
  int diag_only = 0;  // internal (edge pass): assemble_gram_kernel produces only the tiles touching the padding
};

// Relative error bound of a kernel entry under the Gram formulation: |d log k / d r^2| * KA * eps * 2*max|x|^2.
// True when it stays below 2e-13 (the parity tolerance on entries is 1e-12).
bool gram_is_accurate(const CovParams &cp, int dp, double max_sq_norm);

int32_t assemble(Ctx *ctx, cudaStream_t st, const CovParams &cp, const AssembleArgs &a);
}  // namespace dgp
