// solve.cu -- the O(N^2) tail of the path: blocked triangular solves built on the diagonal-block
// inverses, deterministic reductions, GEMV and layout helpers.  All HBM-read bound.
#include "solve.cuh"

namespace dgp {

namespace {

constexpr int SB = 128;  // block edge (= leaf block)

constexpr int ST = 1024;  // threads per CTA of the solve steps

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// Pull a 128 x 128 block (column-major, leading dimension ld) towards L2: 8 lines of 128 B per
// column, one line per thread of a 1024-thread CTA.  The solve steps are latency bound (one dependent
// HBM round trip per launch), so each launch prefetches what the *next* launch will read.
__device__ __forceinline__ void prefetch_block_l2(const double *M, int64_t ld) {
  const int c = threadIdx.x;
  prefetch_l2(M + (c & 7) * 16 + (int64_t)(c >> 3) * ld);
}

// sum_{j<16} M[r + j*ld] * v[j]: one eighth of a 128-wide block row; the 16 loads are independent
// and the compiler keeps them all in flight (1024 threads per CTA give the memory-level parallelism
// that a wider per-thread loop does not: ptxas interleaves loads and FMAs to save registers).
__device__ __forceinline__ double rowpart_mv(const double *M, int64_t ld, const double *v, int r) {
  double m[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) m[u] = M[r + (int64_t)u * ld];
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
  for (int u = 0; u < 16; u += 4) {
    acc0 = fma(m[u + 0], v[u + 0], acc0);
    acc1 = fma(m[u + 1], v[u + 1], acc1);
    acc2 = fma(m[u + 2], v[u + 2], acc2);
    acc3 = fma(m[u + 3], v[u + 3], acc3);
  }
  return (acc0 + acc1) + (acc2 + acc3);
}

// 128-row block times a 128-vector with 1024 threads: thread (r, h) does columns [16h, 16h+16) of
// row r; the eight parts meet in `part` (8 x 128).  Returns the full sum to the h == 0 threads.
__device__ __forceinline__ double rowblock_mv(const double *M, int64_t ld, const double *v, double *part) {
  const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
  part[h * SB + r] = rowpart_mv(M + (int64_t)(h * 16) * ld, ld, v + h * 16, r);
  __syncthreads();
  double s = 0.0;
  if (h == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) s += part[q * SB + r];
  }
  return s;
}

// out[c] = sum_{k<128} M[k + c*ld] * v[k] for c = 0..127: 32 warps x 4 columns, lanes over k
// (16 loads per lane in flight).
template <typename Store>
__device__ __forceinline__ void colblock_dot(const double *M, int64_t ld, const double *v, Store store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double v0 = v[lane], v1 = v[lane + 32], v2 = v[lane + 64], v3 = v[lane + 96];
  double m[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const double *col = M + (int64_t)(warp * 4 + u) * ld;
    m[u][0] = col[lane];
    m[u][1] = col[lane + 32];
    m[u][2] = col[lane + 64];
    m[u][3] = col[lane + 96];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    double s = (m[u][0] * v0 + m[u][1] * v1) + (m[u][2] * v2 + m[u][3] * v3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) store(warp * 4 + u, s);
  }
}

// Forward step b: every row block i > b subtracts L(i,b) z_b; the CTA of block b+1 then applies
// inv(L_{b+1,b+1}) so that the next launch finds z_{b+1} ready.  b == -1: only z_0 = inv_0 x_0.
__global__ void __launch_bounds__(ST) trsv_fwd_step_kernel(const double *L, int64_t ldl, const double *inv, double *x,
                                                           int b) {
  __shared__ double v[SB];
  __shared__ double part[8 * SB];
  const int tid = threadIdx.x, r = tid & 127, h = tid >> 7;
  const int i = b + 1 + blockIdx.x;
  // next launch: this row block reads L(i, b+1) (or, as the new head block, its diagonal inverse)
  if (i > b + 1) prefetch_block_l2(L + (int64_t)i * SB + (int64_t)(b + 1) * SB * ldl, ldl);
  if (i <= b + 2) prefetch_block_l2(inv + (int64_t)i * SB * SB, SB);
  double xi = x[(int64_t)i * SB + r];
  if (b >= 0) {
    if (h == 0) v[r] = x[(int64_t)b * SB + r];
    __syncthreads();
    xi -= rowblock_mv(L + (int64_t)i * SB + (int64_t)b * SB * ldl, ldl, v, part);
  }
  if (i == b + 1) {
    __syncthreads();
    if (h == 0) v[r] = xi;
    __syncthreads();
    xi = rowblock_mv(inv + (int64_t)i * SB * SB, SB, v, part);  // lower triangular with explicit zeros
  }
  if (h == 0) x[(int64_t)i * SB + r] = xi;
}

// Backward step b: column blocks j < b subtract L(b,j)^T x_b; the CTA of block b-1 then applies
// inv(L_{b-1,b-1})^T.  b == nblk: only x_{nblk-1} = inv^T z_{nblk-1}.
__global__ void __launch_bounds__(ST) trsv_bwd_step_kernel(const double *L, int64_t ldl, const double *inv, double *x,
                                                           int b, int nblk) {
  __shared__ double v[SB];
  __shared__ double w[SB];
  const int tid = threadIdx.x;
  const int j = (b == nblk) ? nblk - 1 : (int)blockIdx.x;
  // next launch (step b-1): this column block reads L(b-1, j), or its own diagonal inverse as the head
  if (b >= 1) {
    if (j < b - 1) prefetch_block_l2(L + (int64_t)(b - 1) * SB + (int64_t)j * SB * ldl, ldl);
    else if (j == b - 1) prefetch_block_l2(inv + (int64_t)j * SB * SB, SB);
  }
  if (tid < SB) w[tid] = x[(int64_t)j * SB + tid];
  if (b < nblk) {
    if (tid < SB) v[tid] = x[(int64_t)b * SB + tid];
    __syncthreads();
    colblock_dot(L + (int64_t)b * SB + (int64_t)j * SB * ldl, ldl, v, [&](int c, double s) { w[c] -= s; });
  }
  __syncthreads();
  if (j == b - 1) {
    if (tid < SB) v[tid] = w[tid];
    __syncthreads();
    colblock_dot(inv + (int64_t)j * SB * SB, SB, v, [&](int c, double s) { w[c] = s; });
    __syncthreads();
  }
  if (tid < SB) x[(int64_t)j * SB + tid] = w[tid];
}

// ---- chained solves: ONE launch per direction ------------------------------------------------------------------
// One CTA per 128-row block, all in one grid.  CTA i consumes the solved segments x_j as their CTAs publish them
// (a release store of the call's epoch to flags[j]; the consumer polls with an acquire load), so the 2 N/128
// dependent launches of the step kernels above -- 8-15 us each, launch latency included -- become one dependent
// flag hand-over (~1-2 us) per block, and the off-diagonal blocks of a row stream while earlier segments are
// still being solved.  A CTA waits only for CTAs of lower blockIdx; the hardware dispatches CTAs in index order,
// so every CTA waited for is resident or finished, whatever else occupies the GPU.
// The arithmetic is the step kernels', operation for operation: results are bit-identical.
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

constexpr size_t CHAIN_SMEM = (size_t)(SB * SB + 2 * 8 * SB + SB) * sizeof(double);

__global__ void __launch_bounds__(ST, 1) trsv_fwd_chain_kernel(const double *L, int64_t ldl, const double *inv, double *x,
                                                               int *flags, int epoch) {
  extern __shared__ __align__(16) double sm[];
  double *invs = sm;                 // inverse of this block's diagonal block, staged while the row streams
  double *part = sm + SB * SB;       // [2][8][SB]
  double *v = part + 2 * 8 * SB;     // [SB]
  const int tid = threadIdx.x, r = tid & 127, h = tid >> 7;
  const int i = blockIdx.x;
  {
    const double *src = inv + (int64_t)i * SB * SB;
#pragma unroll
    for (int u = 0; u < 16; ++u) invs[tid + u * ST] = src[tid + u * ST];
  }
  double xi = (h == 0) ? x[(int64_t)i * SB + r] : 0.0;
  const double *Lrow = L + (int64_t)i * SB + r + (int64_t)(h * 16) * ldl;
  for (int j = 0; j < i; ++j) {
    double m[16];
    const double *blk = Lrow + (int64_t)j * SB * ldl;
#pragma unroll
    for (int u = 0; u < 16; ++u) m[u] = blk[(int64_t)u * ldl];   // in flight while the flag is polled
    if (tid == 0)
      while (ld_acquire_gpu(flags + j) != epoch) {}
    __syncthreads();   // x_j is visible; the parts of block j-1 are complete
    if (h == 0 && j > 0) {
      const double *pp = part + ((j - 1) & 1) * 8 * SB;
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += pp[q * SB + r];
      xi -= s;
    }
    const double *xj = x + (int64_t)j * SB + h * 16;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll
    for (int u = 0; u < 16; u += 4) {
      acc0 = fma(m[u + 0], __ldcg(xj + u + 0), acc0);
      acc1 = fma(m[u + 1], __ldcg(xj + u + 1), acc1);
      acc2 = fma(m[u + 2], __ldcg(xj + u + 2), acc2);
      acc3 = fma(m[u + 3], __ldcg(xj + u + 3), acc3);
    }
    part[(j & 1) * 8 * SB + h * SB + r] = (acc0 + acc1) + (acc2 + acc3);
  }
  __syncthreads();
  if (h == 0) {
    if (i > 0) {
      const double *pp = part + ((i - 1) & 1) * 8 * SB;
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += pp[q * SB + r];
      xi -= s;
    }
    v[r] = xi;
  }
  __syncthreads();
  xi = rowblock_mv(invs, SB, v, part);  // lower triangular with explicit zeros
  if (h == 0) {
    x[(int64_t)i * SB + r] = xi;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) st_release_gpu(flags + i, epoch);
}

__global__ void __launch_bounds__(ST, 1) trsv_bwd_chain_kernel(const double *L, int64_t ldl, const double *inv, double *x,
                                                               int *flags, int epoch, int nblk) {
  extern __shared__ __align__(16) double sm[];
  double *invs = sm;
  double *v = sm + SB * SB;   // [SB]  x_b of the block being consumed
  double *w = v + SB;         // [SB]  this block's running right-hand side
  const int tid = threadIdx.x;
  const int j = nblk - 1 - (int)blockIdx.x;   // column block solved here; waits for blocks b > j = lower blockIdx
  {
    const double *src = inv + (int64_t)j * SB * SB;
#pragma unroll
    for (int u = 0; u < 16; ++u) invs[tid + u * ST] = src[tid + u * ST];
  }
  if (tid < SB) w[tid] = x[(int64_t)j * SB + tid];
  const int warp = tid >> 5, lane = tid & 31;
  for (int b = nblk - 1; b > j; --b) {
    double m[4][4];
    const double *blk = L + (int64_t)b * SB + (int64_t)j * SB * ldl;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double *col = blk + (int64_t)(warp * 4 + u) * ldl;
      m[u][0] = col[lane];
      m[u][1] = col[lane + 32];
      m[u][2] = col[lane + 64];
      m[u][3] = col[lane + 96];
    }
    if (tid == 0)
      while (ld_acquire_gpu(flags + (nblk - 1 - b)) != epoch) {}
    __syncthreads();
    const double *xb = x + (int64_t)b * SB;
    const double v0 = __ldcg(xb + lane), v1 = __ldcg(xb + lane + 32), v2 = __ldcg(xb + lane + 64),
                 v3 = __ldcg(xb + lane + 96);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double s = (m[u][0] * v0 + m[u][1] * v1) + (m[u][2] * v2 + m[u][3] * v3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) w[warp * 4 + u] -= s;   // each warp owns its four columns of w
    }
  }
  __syncthreads();
  if (tid < SB) v[tid] = w[tid];
  __syncthreads();
  colblock_dot(invs, SB, v, [&](int c, double s) { w[c] = s; });
  __syncthreads();
  if (tid < SB) {
    x[(int64_t)j * SB + tid] = w[tid];
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) st_release_gpu(flags + blockIdx.x, epoch);
}

__device__ __forceinline__ double block_reduce_1024(double s, double *sh) {
  const int tid = threadIdx.x;
  sh[tid] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (tid < o) sh[tid] += sh[tid + o];
    __syncthreads();
  }
  return sh[0];
}

__global__ void __launch_bounds__(1024) diag_log_sum_kernel(const double *A, int64_t lda, int64_t n, double *out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += log(A[i + i * lda]);
  s = block_reduce_1024(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(1024) dot_kernel(const double *x, const double *y, int64_t n, double *out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fma(x[i], y[i], s);
  s = block_reduce_1024(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void __launch_bounds__(256) gemv_t_kernel(const double *A, int64_t lda, int64_t rows, const double *x,
                                                     const double *base, double *y) {
  __shared__ double sh[256];
  const int64_t j = blockIdx.x;
  const double *col = A + j * lda;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) s = fma(col[i], x[i], s);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) y[j] = (base ? base[j] : 0.0) + sh[0];
}

__global__ void __launch_bounds__(128) lower_row_sums_kernel(const double *L, int64_t ldl, int64_t n, double *bar) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int64_t j = 0; j <= i; ++j) s += L[i + j * ldl];
  bar[i] = s;
}

__global__ void __launch_bounds__(256) mirror_lower_kernel(double *A, int64_t lda) {
  __shared__ double tile[32][33];
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (tj > ti) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) tile[r][tx] = A[(int64_t)ti * 32 + tx + ((int64_t)tj * 32 + r) * lda];  // tile[col][row]
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // destination element (row = tj*32 + tx, col = ti*32 + r) = source (row = ti*32 + r, col = tj*32 + tx)
    if (ti != tj || tx < r) A[(int64_t)tj * 32 + tx + ((int64_t)ti * 32 + r) * lda] = tile[tx][r];
  }
}

}  // namespace

// Flag words of the chained solves: one array per stream (solves on one stream are ordered, so they can share it;
// every call uses a fresh epoch value, so the words never need resetting).
static int32_t chain_flags(Ctx *ctx, cudaStream_t st, int nblk, int **flags, int *epoch) {
  char name[48];
  snprintf(name, sizeof name, "trsv_flags@%p", (void *)st);
  const bool fresh = ctx->arena.find(name) == ctx->arena.end() || ctx->arena[name].second < (size_t)2 * nblk * sizeof(int);
  int *f = (int *)ctx->buf(name, (size_t)2 * nblk * sizeof(int));
  if (!f) return DGP_ERR_OOM;
  if (fresh) DGP_CUDA_OK(ctx, cudaMemsetAsync(f, 0, (size_t)2 * nblk * sizeof(int), st));
  *flags = f;
  *epoch = ++ctx->trsv_epoch;
  return DGP_OK;
}

int32_t solve_init(Ctx *ctx) {
  DGP_CUDA_OK(ctx, cudaFuncSetAttribute(trsv_fwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM));
  DGP_CUDA_OK(ctx, cudaFuncSetAttribute(trsv_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHAIN_SMEM));
  return DGP_OK;
}

int32_t trsv_lower_fwd(Ctx *ctx, cudaStream_t st, const double *L, int64_t ldl, const double *inv, int64_t n,
                       double *x) {
  const int nblk = (int)(n / SB);
  if (nblk == 0) return DGP_OK;
  if (ctx->trsv_chained) {
    int *flags = nullptr, epoch = 0;
    DGP_TRY(chain_flags(ctx, st, nblk, &flags, &epoch));
    trsv_fwd_chain_kernel<<<nblk, ST, CHAIN_SMEM, st>>>(L, ldl, inv, x, flags, epoch);
    ctx->launches++;
  } else {
    trsv_fwd_step_kernel<<<1, ST, 0, st>>>(L, ldl, inv, x, -1);
    for (int b = 0; b + 1 < nblk; ++b) trsv_fwd_step_kernel<<<nblk - 1 - b, ST, 0, st>>>(L, ldl, inv, x, b);
    ctx->launches += nblk;
  }
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t trsv_lower_bwd(Ctx *ctx, cudaStream_t st, const double *L, int64_t ldl, const double *inv, int64_t n,
                       double *x) {
  const int nblk = (int)(n / SB);
  if (nblk == 0) return DGP_OK;
  if (ctx->trsv_chained) {
    int *flags = nullptr, epoch = 0;
    DGP_TRY(chain_flags(ctx, st, nblk, &flags, &epoch));
    trsv_bwd_chain_kernel<<<nblk, ST, CHAIN_SMEM, st>>>(L, ldl, inv, x, flags + nblk, epoch, nblk);
    ctx->launches++;
  } else {
    trsv_bwd_step_kernel<<<1, ST, 0, st>>>(L, ldl, inv, x, nblk, nblk);
    for (int b = nblk - 1; b >= 1; --b) trsv_bwd_step_kernel<<<b, ST, 0, st>>>(L, ldl, inv, x, b, nblk);
    ctx->launches += nblk;
  }
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t diag_log_sum(Ctx *ctx, cudaStream_t st, const double *A, int64_t lda, int64_t n, double *out) {
  diag_log_sum_kernel<<<1, 1024, 0, st>>>(A, lda, n, out);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t dot(Ctx *ctx, cudaStream_t st, const double *x, const double *y, int64_t n, double *out) {
  dot_kernel<<<1, 1024, 0, st>>>(x, y, n, out);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t gemv_t(Ctx *ctx, cudaStream_t st, const double *A, int64_t lda, int64_t rows, int64_t cols, const double *x,
               const double *base, double *y) {
  if (cols == 0) return DGP_OK;
  gemv_t_kernel<<<(unsigned)cols, 256, 0, st>>>(A, lda, rows, x, base, y);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t lower_row_sums(Ctx *ctx, cudaStream_t st, const double *L, int64_t ldl, int64_t n, double *bar) {
  if (n == 0) return DGP_OK;
  lower_row_sums_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(L, ldl, n, bar);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t mirror_lower(Ctx *ctx, cudaStream_t st, double *A, int64_t lda, int64_t n) {
  if (n == 0) return DGP_OK;
  if (n % 32) return ctx->fail(DGP_ERR_SHAPE, "mirror_lower: n must be a multiple of 32");
  dim3 grid((unsigned)(n / 32), (unsigned)(n / 32));
  mirror_lower_kernel<<<grid, 256, 0, st>>>(A, lda);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

int32_t copy_block(Ctx *ctx, cudaStream_t st, const double *src, int64_t lds, double *dst, int64_t ldd, int64_t n,
                   int64_t m) {
  if (n == 0 || m == 0) return DGP_OK;
  DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), n * sizeof(double), m,
                                     cudaMemcpyDeviceToDevice, st));
  return DGP_OK;
}

}  // namespace dgp
