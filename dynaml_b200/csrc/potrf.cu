// potrf.cu -- right-looking blocked Cholesky with look-ahead, the triangular inverse and
// K^-1 = L^-T L^-1, all driven through the DMMA GEMM of gemm.cu.
//
// Reference semantics (CORE/algebra/bcholesky.scala:85-125): cholesky(A11); L21 = A21 inv(L11)^T;
// recurse on A22 - L21 L21^T.  The reference multiplies by an explicit inverse of the diagonal
// block, and so do we: the 128x128 leaf kernel returns both L11 and inv(L11), which turns every
// triangular solve of the path into a tensor-core GEMM.
#include "potrf.cuh"
#include "gemm.cuh"

namespace dgp {

namespace {

constexpr int LB = 128;          // leaf block
constexpr int LEAF_THREADS = 256;
constexpr size_t LEAF_SMEM = 0;

// Cholesky + triangular inverse of one 128x128 diagonal block, register resident.
//
// The 256 threads form a 16x16 grid (tr, tc); thread (tr, tc) owns the elements (r, c) with
// r = 16 i + tr, c = 16 q + tc (2D block-cyclic), i.e. an 8x8 local block of which only the local
// lower part i >= q is live (36 FP64 values for L, 36 for W = L^-1).  Each elimination step
// broadcasts one column of L and one row of W through double-buffered 128-entry lines in shared
// memory and costs a single __syncthreads; all register indices are compile-time after unrolling
// the 16-wide column blocks, so nothing spills.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_inv_kernel(double *A, int64_t lda, double *Inv, int *info, int pivot_base) {
  __shared__ double colbuf[2][LB];
  __shared__ double rowbuf[2][LB];
  const int tid = threadIdx.x;
  const int tr = tid & 15, tc = tid >> 4;

  double a[8][8], w[8][8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      w[i][q] = 0.0;
      a[i][q] = (i >= q) ? A[(i * 16 + tr) + (int64_t)(q * 16 + tc) * lda] : 0.0;
    }

  // ---- factorisation and inverse, fused: one elimination step and one __syncthreads per column ----
  // Step k: column k of the trailing matrix and row k of R = (running) L W - I travel through shared
  // memory; then  l = sqrt(a_kk), L(:,k) = a(:,k)/l, a(r,c) -= L(r,k) L(c,k)   (Cholesky)
  // and           W(k,:) = R(k,:)/l,  R(i,:) -= L(i,k) W(k,:) for i > k        (L W = I, right-looking).
#pragma unroll
  for (int kc = 0; kc < 8; ++kc) {
    for (int kk = 0; kk < 16; ++kk) {
      const int k = kc * 16 + kk;
      double *cb = colbuf[k & 1];
      double *rb = rowbuf[k & 1];
      if (tc == kk) {
#pragma unroll
        for (int i = kc; i < 8; ++i) cb[i * 16 + tr] = a[i][kc];
      }
      if (tr == kk) {
#pragma unroll
        for (int q = 0; q <= kc; ++q) rb[q * 16 + tc] = w[kc][q];
      }
      __syncthreads();
      double d = cb[k];
      if (!(d > 0.0)) {  // also catches NaN; uniform across the CTA
        if (tid == 0) atomicCAS(info, 0, pivot_base + k + 1);
        d = 1.0;
      }
      // every thread needs 1/l: rsqrt + one multiply instead of a redundant sqrt and divide on all
      // 256 threads (the FP64 pipe, not latency, bounds this step); l = d * rsqrt(d) is within 1 ulp
      const double li = rsqrt(d);
      const double l = d * li;
      double lr[8], lc[8], wk[8];
#pragma unroll
      for (int i = kc; i < 8; ++i) {
        lr[i] = cb[i * 16 + tr] * li;  // L(r, k) for this thread's rows
        lc[i] = cb[i * 16 + tc] * li;  // L(c, k) for this thread's columns
      }
#pragma unroll
      for (int q = 0; q <= kc; ++q) {
        const int c = q * 16 + tc;
        wk[q] = (c == k) ? li : ((c < k) ? rb[c] * li : 0.0);  // W(k, c)
      }
      if (tc == kk) {  // owners of column k keep the finished column of L
#pragma unroll
        for (int i = kc; i < 8; ++i) {
          const int r = i * 16 + tr;
          if (r > k) a[i][kc] = lr[i];
          else if (r == k) a[i][kc] = l;
        }
      }
      if (tr == kk) {  // owners of row k keep the finished row of W
#pragma unroll
        for (int q = 0; q <= kc; ++q) w[kc][q] = wk[q];
      }
      // Cholesky trailing update on the live local part (rows > k, columns > k)
#pragma unroll
      for (int q = kc; q < 8; ++q) {
        const bool cq = (q > kc) || (tc > kk);
#pragma unroll
        for (int i = q; i < 8; ++i) {
          const bool ri = (i > kc) || (tr > kk);
          if (cq && ri) a[i][q] = fma(-lr[i], lc[q], a[i][q]);
        }
      }
      // inverse update: rows > k, columns <= k (wk is zero beyond column k)
#pragma unroll
      for (int q = 0; q <= kc; ++q) {
#pragma unroll
        for (int i = kc; i < 8; ++i) {
          if (i >= q) {
            const bool ri = (i > kc) || (tr > kk);
            if (ri) w[i][q] = fma(-lr[i], wk[q], w[i][q]);
          }
        }
      }
    }
  }

  // ---- write back: L (upper part zeroed) and W --------------------------------------------------
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 16 + tr, c = q * 16 + tc;
      double lv = 0.0, wv = 0.0;
      if (i >= q && r >= c) {
        lv = a[i][q];
        wv = w[i][q];
      }
      A[r + (int64_t)c * lda] = lv;
      Inv[r + c * LB] = wv;
    }
}

// W diagonal tiles <- leaf inverses
__global__ void copy_diag_blocks_kernel(const double *inv, double *W, int64_t ldw) {
  const int b = blockIdx.x;
  const double *src = inv + (int64_t)b * LB * LB;
  double *dst = W + (int64_t)b * LB * (ldw + 1);
  for (int idx = threadIdx.x; idx < LB * LB; idx += blockDim.x) {
    int r = idx & (LB - 1), c = idx >> 7;
    dst[r + (int64_t)c * ldw] = src[idx];
  }
}

// U diagonal tiles <- transposed leaf inverses
__global__ void copy_diag_blocks_transposed_kernel(const double *inv, double *U, int64_t ldu) {
  __shared__ double tile[32][33];
  const int b = blockIdx.x;
  const double *src = inv + (int64_t)b * LB * LB;
  double *dst = U + (int64_t)b * LB * (ldu + 1);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int bt = 0; bt < 16; ++bt) {
    const int r0 = (bt & 3) * 32, c0 = (bt >> 2) * 32;
    for (int j = ty; j < 32; j += 8) tile[j][tx] = src[r0 + tx + (c0 + j) * LB];  // tile[c][r] = inv(r0 + r, c0 + c)
    __syncthreads();
    for (int j = ty; j < 32; j += 8) dst[c0 + tx + (int64_t)(r0 + j) * ldu] = tile[tx][j];  // U(c0 + c, r0 + r)
    __syncthreads();
  }
}

cudaEvent_t get_event(Ctx *ctx, size_t i) {
  while (ctx->ev_pool.size() <= i) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    ctx->ev_pool.push_back(e);
  }
  return ctx->ev_pool[i];
}

// Factor the block column starting at c0 (width w), 128 columns at a time, left-looking inside the
// block column: before a 128-wide sub-panel is factored it receives, in ONE product with K = cs - c0,
// the updates of all sub-panels already finished in this block column (the right-looking variant
// issued a rank-128 update per sub-panel; profiles/launches_c3_step_r1.txt showed those K = 128
// launches at half the efficiency of the K >= 256 ones).  Enqueued on `st`.
int32_t factor_block_column(Ctx *ctx, cudaStream_t st, double *A, int64_t lda, int64_t n, int64_t c0, int64_t w,
                            double *inv, int *info, double *scratch) {
  for (int64_t cs = c0; cs < c0 + w; cs += LB) {
    if (cs > c0) {
      GemmArgs u;  // A[cs:, cs:cs+128] -= A[cs:, c0:cs] * A[cs:cs+128, c0:cs]^T
      u.A = A + cs + c0 * lda;  u.lda = lda;  u.a_kmajor = 0;
      u.B = A + cs + c0 * lda;  u.ldb = lda;  u.b_kmajor = 0;   // op(B)(k,n) = L(cs + n, c0 + k)
      u.C = A + cs + cs * lda;  u.ldc = lda;
      u.M = (int)(n - cs);  u.N = LB;  u.K = (int)(cs - c0);
      u.alpha = -1.0;  u.beta = 1.0;
      DGP_TRY(gemm(ctx, st, u));
    }
    double *invb = inv + (cs / LB) * LB * LB;
    leaf_potrf_inv_kernel<<<1, LEAF_THREADS, LEAF_SMEM, st>>>(A + cs + cs * lda, lda, invb, info, (int)cs);
    ctx->launches++;
    DGP_CUDA_OK(ctx, cudaGetLastError());
    const int64_t rb = cs + LB;
    if (rb >= n) break;
    // panel: A[rb:, cs:cs+128] <- A[rb:, cs:cs+128] * inv(L_cs)^T, through a scratch copy of the
    // panel (two CTAs share each 128-row strip, so the product cannot run in place)
    const int64_t mrows = n - rb;
    DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(scratch, mrows * sizeof(double), A + rb + cs * lda, lda * sizeof(double),
                                       mrows * sizeof(double), LB, cudaMemcpyDeviceToDevice, st));
    GemmArgs g;
    g.A = scratch;            g.lda = mrows;  g.a_kmajor = 0;
    g.B = invb;               g.ldb = LB;     g.b_kmajor = 0;   // op(B)(k,n) = inv(n,k)
    g.C = A + rb + cs * lda;  g.ldc = lda;
    g.M = (int)mrows;  g.N = LB;  g.K = LB;
    g.alpha = 1.0;  g.beta = 0.0;
    DGP_TRY(gemm(ctx, st, g));
  }
  return DGP_OK;
}

}  // namespace

int32_t potrf_init(Ctx *ctx) {
  (void)ctx;
  return DGP_OK;
}

// Wider blocks amortise the per-tile prologue/epilogue of the trailing update (measured on B200,
// scripts/tune.py: N = 32768 -> 404 / 386 / 382 ms at nb = 256 / 512 / 768), narrower ones shorten
// the leaf -> panel -> update chain that bounds small N.
int64_t potrf_outer_block(const Ctx *ctx, int64_t n) {
  return ctx->nb > 0 ? ctx->nb : (n >= 24576 ? 768 : (n >= 12288 ? 512 : 256));
}

int32_t potrf_blocked(Ctx *ctx, double *A, int64_t lda, int64_t n, double *inv, int *info) {
  const bool la = ctx->lookahead && n > 256;
  return potrf_blocked_on(ctx, ctx->stream, la ? ctx->panel : ctx->stream, A, lda, n, inv, info, "potrf_panel");
}

int32_t potrf_blocked_on(Ctx *ctx, cudaStream_t S, cudaStream_t P, double *A, int64_t lda, int64_t n, double *inv,
                         int *info, const char *scratch_name) {
  if (n % LB) return ctx->fail(DGP_ERR_SHAPE, "potrf: n=%lld must be a multiple of 128", (long long)n);
  const int64_t nb = potrf_outer_block(ctx, n);
  const bool la = (P != S) && n > nb;
  if (!la) P = S;
  size_t ev = 0;

  double *scratch = (double *)ctx->buf(scratch_name, (size_t)n * LB * sizeof(double));
  if (!scratch) return DGP_ERR_OOM;
  if (la) {  // the panel stream starts after whatever produced A on the main stream
    cudaEvent_t e = get_event(ctx, ev++);
    DGP_CUDA_OK(ctx, cudaEventRecord(e, S));
    DGP_CUDA_OK(ctx, cudaStreamWaitEvent(P, e, 0));
  }

  for (int64_t c0 = 0; c0 < n; c0 += nb) {
    const int64_t w = (n - c0 < nb) ? (n - c0) : nb;
    DGP_TRY(factor_block_column(ctx, P, A, lda, n, c0, w, inv, info, scratch));
    const int64_t c1 = c0 + w;
    if (c1 >= n) break;
    if (la) {
      cudaEvent_t e = get_event(ctx, ev++);
      DGP_CUDA_OK(ctx, cudaEventRecord(e, P));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(S, e, 0));
    }
    const int64_t w1 = (n - c1 < nb) ? (n - c1) : nb;
    // (a) trailing update restricted to the next block column
    GemmArgs a;
    a.A = A + c1 + c0 * lda;  a.lda = lda;  a.a_kmajor = 0;
    a.B = A + c1 + c0 * lda;  a.ldb = lda;  a.b_kmajor = 0;
    a.C = A + c1 + c1 * lda;  a.ldc = lda;
    a.M = (int)(n - c1);  a.N = (int)w1;  a.K = (int)w;
    a.alpha = -1.0;  a.beta = 1.0;
    DGP_TRY(gemm(ctx, S, a));
    if (la) {
      cudaEvent_t e = get_event(ctx, ev++);
      DGP_CUDA_OK(ctx, cudaEventRecord(e, S));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(P, e, 0));
    }
    // (b) the rest of the trailing matrix (lower tiles), overlapped with the next panel
    const int64_t c2 = c1 + w1;
    if (c2 < n) {
      GemmArgs b;
      b.A = A + c2 + c0 * lda;  b.lda = lda;  b.a_kmajor = 0;
      b.B = A + c2 + c0 * lda;  b.ldb = lda;  b.b_kmajor = 0;
      b.C = A + c2 + c2 * lda;  b.ldc = lda;
      b.M = (int)(n - c2);  b.N = (int)(n - c2);  b.K = (int)w;
      b.alpha = -1.0;  b.beta = 1.0;
      b.lower_only = 1;
      DGP_TRY(gemm(ctx, S, b));
    }
  }
  if (la) {  // join
    cudaEvent_t e = get_event(ctx, ev++);
    DGP_CUDA_OK(ctx, cudaEventRecord(e, P));
    DGP_CUDA_OK(ctx, cudaStreamWaitEvent(S, e, 0));
  }
  return DGP_OK;
}

int32_t trtri_lower(Ctx *ctx, const double *L, int64_t ldl, const double *inv, int64_t n, double *W, int64_t ldw,
                    double *scratch, int64_t lds) {
  return trtri_lower_on(ctx, ctx->stream, L, ldl, inv, n, W, ldw, scratch, lds);
}

int32_t trtri_lower_on(Ctx *ctx, cudaStream_t S, const double *L, int64_t ldl, const double *inv, int64_t n, double *W,
                       int64_t ldw, double *scratch, int64_t lds) {
  if (n % LB) return ctx->fail(DGP_ERR_SHAPE, "trtri: n must be a multiple of 128");
  copy_diag_blocks_kernel<<<(unsigned)(n / LB), 256, 0, S>>>(inv, W, ldw);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  // Level h joins the inverses of two adjacent diagonal blocks [r0, r0+h) and [r0+h, r0+h+h2):
  //   W21 = -W22 * (L21 * W11)
  for (int64_t h = LB; h < n; h *= 2) {
    const int64_t nfull = n / (2 * h);             // problems with h2 == h
    const int64_t rem = n - nfull * 2 * h;         // leftover rows at the end
    const int64_t h2_last = rem > h ? rem - h : 0; // a last, smaller problem
    for (int pass = 0; pass < 2; ++pass) {
      const int64_t cnt = pass == 0 ? nfull : (h2_last > 0 ? 1 : 0);
      if (cnt == 0) continue;
      const int64_t r0 = pass == 0 ? 0 : nfull * 2 * h;
      const int64_t h2 = pass == 0 ? h : h2_last;
      GemmArgs t;  // T = L21 * W11  (W11 lower triangular: k >= n0)
      t.A = L + (r0 + h) + r0 * ldl;  t.lda = ldl;  t.a_kmajor = 0;
      t.B = W + r0 + r0 * ldw;        t.ldb = ldw;  t.b_kmajor = 1;
      t.C = scratch + (r0 + h) + r0 * lds;  t.ldc = lds;
      t.M = (int)h2;  t.N = (int)h;  t.K = (int)h;
      t.alpha = 1.0;  t.beta = 0.0;
      t.kmode = K_GE_COL;
      t.batch = (int)cnt;
      t.strideA = 2 * h * (ldl + 1);  t.strideB = 2 * h * (ldw + 1);  t.strideC = 2 * h * (lds + 1);
      DGP_TRY(gemm(ctx, S, t));
      GemmArgs u;  // W21 = -W22 * T  (W22 lower triangular: k < m0 + 128)
      u.A = W + (r0 + h) + (r0 + h) * ldw;  u.lda = ldw;  u.a_kmajor = 0;
      u.B = scratch + (r0 + h) + r0 * lds;  u.ldb = lds;  u.b_kmajor = 1;
      u.C = W + (r0 + h) + r0 * ldw;        u.ldc = ldw;
      u.M = (int)h2;  u.N = (int)h;  u.K = (int)h2;
      u.alpha = -1.0;  u.beta = 0.0;
      u.kmode = K_LT_ROWEND;
      u.reverse = 1;
      u.batch = (int)cnt;
      u.strideA = 2 * h * (ldw + 1);  u.strideB = 2 * h * (lds + 1);  u.strideC = 2 * h * (ldw + 1);
      DGP_TRY(gemm(ctx, S, u));
    }
  }
  return DGP_OK;
}

int32_t trtri_lower_transposed_on(Ctx *ctx, cudaStream_t S, const double *L, int64_t ldl, const double *inv, int64_t n,
                                  double *U, int64_t ldu, double *scratch, int64_t lds) {
  if (n % LB) return ctx->fail(DGP_ERR_SHAPE, "trtri: n must be a multiple of 128");
  copy_diag_blocks_transposed_kernel<<<(unsigned)(n / LB), 256, 0, S>>>(inv, U, ldu);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  // Level h joins the transposed inverses of two adjacent diagonal blocks [r0, r0+h) and [r0+h, r0+h+h2):
  //   U12 = W21^T = -(W22 L21 W11)^T = -(U11 * L21^T) * U22
  for (int64_t h = LB; h < n; h *= 2) {
    const int64_t nfull = n / (2 * h);
    const int64_t rem = n - nfull * 2 * h;
    const int64_t h2_last = rem > h ? rem - h : 0;
    for (int pass = 0; pass < 2; ++pass) {
      const int64_t cnt = pass == 0 ? nfull : (h2_last > 0 ? 1 : 0);
      if (cnt == 0) continue;
      const int64_t r0 = pass == 0 ? 0 : nfull * 2 * h;
      const int64_t h2 = pass == 0 ? h : h2_last;
      GemmArgs t;  // X = U11 * L21^T  (NT; U11 upper triangular: k >= m0)
      t.A = U + r0 + r0 * ldu;        t.lda = ldu;  t.a_kmajor = 0;
      t.B = L + (r0 + h) + r0 * ldl;  t.ldb = ldl;  t.b_kmajor = 0;   // op(B)(k,n) = L21(n,k)
      t.C = scratch + r0 + (r0 + h) * lds;  t.ldc = lds;
      t.M = (int)h;  t.N = (int)h2;  t.K = (int)h;
      t.alpha = 1.0;  t.beta = 0.0;
      t.kmode = K_GE_ROW;
      t.batch = (int)cnt;
      t.strideA = 2 * h * (ldu + 1);  t.strideB = 2 * h * (ldl + 1);  t.strideC = 2 * h * (lds + 1);
      DGP_TRY(gemm(ctx, S, t));
      GemmArgs u;  // U12 = -X * U22  (NN; U22 upper triangular: k < n0 + 64)
      u.A = scratch + r0 + (r0 + h) * lds;  u.lda = lds;  u.a_kmajor = 0;
      u.B = U + (r0 + h) + (r0 + h) * ldu;  u.ldb = ldu;  u.b_kmajor = 1;
      u.C = U + r0 + (r0 + h) * ldu;        u.ldc = ldu;
      u.M = (int)h;  u.N = (int)h2;  u.K = (int)h2;
      u.alpha = -1.0;  u.beta = 0.0;
      u.kmode = K_LT_COLEND;
      u.batch = (int)cnt;
      u.strideA = 2 * h * (lds + 1);  u.strideB = 2 * h * (ldu + 1);  u.strideC = 2 * h * (ldu + 1);
      DGP_TRY(gemm(ctx, S, u));
    }
  }
  return DGP_OK;
}

int32_t trsm_lower_fwd(Ctx *ctx, cudaStream_t S, const double *L, int64_t ldl, const double *inv, int64_t n,
                       double *B, int64_t ldb, int64_t m, int64_t nb) {
  if (n % LB || nb % LB || m % 64) return ctx->fail(DGP_ERR_SHAPE, "trsm: n, nb must be multiples of 128, m of 64");
  for (int64_t c0 = 0; c0 < n; c0 += nb) {
    const int64_t c1 = (c0 + nb < n) ? c0 + nb : n;
    for (int64_t cs = c0; cs < c1; cs += LB) {
      GemmArgs d;  // V_cs = inv(L_cs,cs) * B_cs, in place: a CTA reads exactly the 128 x 64 tile it writes
      d.A = inv + (cs / LB) * LB * LB;  d.lda = LB;  d.a_kmajor = 0;
      d.B = B + cs;  d.ldb = ldb;  d.b_kmajor = 1;
      d.C = B + cs;  d.ldc = ldb;
      d.M = LB;  d.N = (int)m;  d.K = LB;
      DGP_TRY(gemm(ctx, S, d));
      const int64_t rb = cs + LB;
      if (rb >= c1) break;
      GemmArgs u;  // rows of the same block: B[rb:c1] -= L[rb:c1, cs] * V_cs
      u.A = L + rb + cs * ldl;  u.lda = ldl;  u.a_kmajor = 0;
      u.B = B + cs;  u.ldb = ldb;  u.b_kmajor = 1;
      u.C = B + rb;  u.ldc = ldb;
      u.M = (int)(c1 - rb);  u.N = (int)m;  u.K = LB;
      u.alpha = -1.0;  u.beta = 1.0;
      DGP_TRY(gemm(ctx, S, u));
    }
    if (c1 >= n) break;
    GemmArgs t;  // everything below the block: B[c1:] -= L[c1:, c0:c1] * V[c0:c1]
    t.A = L + c1 + c0 * ldl;  t.lda = ldl;  t.a_kmajor = 0;
    t.B = B + c0;  t.ldb = ldb;  t.b_kmajor = 1;
    t.C = B + c1;  t.ldc = ldb;
    t.M = (int)(n - c1);  t.N = (int)m;  t.K = (int)(c1 - c0);
    t.alpha = -1.0;  t.beta = 1.0;
    DGP_TRY(gemm(ctx, S, t));
  }
  return DGP_OK;
}

int32_t lauum_upper_on(Ctx *ctx, cudaStream_t S, const double *U, int64_t ldu, int64_t n, double *Kinv, int64_t ldk) {
  GemmArgs g;  // Kinv(i,j) = sum_{k >= i} U(i,k) U(j,k),  i >= j  (NT)
  g.A = U;  g.lda = ldu;  g.a_kmajor = 0;
  g.B = U;  g.ldb = ldu;  g.b_kmajor = 0;
  g.C = Kinv;  g.ldc = ldk;
  g.M = (int)n;  g.N = (int)n;  g.K = (int)n;
  g.alpha = 1.0;  g.beta = 0.0;
  g.lower_only = 1;
  g.kmode = K_GE_ROW;
  return gemm(ctx, S, g);
}

int32_t lauum_lower(Ctx *ctx, const double *W, int64_t ldw, int64_t n, double *Kinv, int64_t ldk) {
  return lauum_lower_on(ctx, ctx->stream, W, ldw, n, Kinv, ldk);
}

int32_t lauum_lower_on(Ctx *ctx, cudaStream_t S, const double *W, int64_t ldw, int64_t n, double *Kinv, int64_t ldk) {
  GemmArgs g;  // Kinv(i,j) = sum_{k >= i} W(k,i) W(k,j),  i >= j
  g.A = W;  g.lda = ldw;  g.a_kmajor = 1;
  g.B = W;  g.ldb = ldw;  g.b_kmajor = 1;
  g.C = Kinv;  g.ldc = ldk;
  g.M = (int)n;  g.N = (int)n;  g.K = (int)n;
  g.alpha = 1.0;  g.beta = 0.0;
  g.lower_only = 1;
  g.kmode = K_GE_ROW;
  return gemm(ctx, S, g);
}

}  // namespace dgp
