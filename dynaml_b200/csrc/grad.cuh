// grad.cuh -- gradient trace pass interface.
#pragma once
#include "common.cuh"
#include "covfn.cuh"

namespace dgp {

struct GradArgs {
  const double *X = nullptr;      // np x dp row-major points
  const double *Xs = nullptr;     // ARD only: points pre-scaled by 1/b_c
  const double *alpha = nullptr;  // K^-1 r, length np (padding zero)
  const double *Kinv = nullptr;   // lower tiles of K^-1, column-major
  int64_t ld = 0;
  int64_t n = 0, np = 0;
  int dp = 0;
  double *partials = nullptr;     // tiles x ns scratch
  double *sums = nullptr;         // ns moments (device)
  int lean = 0;                   // caller: Gram accuracy gate holds and the points are pairwise distinct
  int edge_only = 0;              // internal: grad_trace_kernel covers only the tiles the lean pass leaves out
};

int grad_partials_per_tile(const CovParams &cp, int dp);
int32_t grad_trace(Ctx *ctx, cudaStream_t st, const CovParams &cp, const GradArgs &a, int *ns);
void grad_finish(const Spec &sp, int d, const double *S, int ns, double *g);
// Moments of single pairs (weight 1): out[(i + (j - c0) * n1) * ns + q] for rows i < n1 of X1 and columns c0 <= j < c0 + ncols
// of X2 (row-major n x dp points; X1s / X2s the ARD-scaled copies, else unused).  grad_finish on a pair's moments gives
// d k(x_i, y_j) / d theta: the gradient matrices of SVMKernel.buildKernelGradMatrix (SVMKernel.scala:107-179).
int32_t grad_pair_moments(Ctx *ctx, cudaStream_t st, const CovParams &cp, const double *X1, const double *X1s,
                          const double *X2, const double *X2s, int64_t n1, int64_t c0, int64_t ncols, int dp, double *out,
                          int *ns_out);

}  // namespace dgp
