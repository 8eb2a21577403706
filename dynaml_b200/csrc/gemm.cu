// gemm.cu -- FP64 GEMM on the sm_100a DMMA tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4;
// tcgen05 has no f64 kind).
//
// One CTA = 256 threads = 8 warps computes a 128x128 tile of C; each warp owns 64x32 of it as
// 8x4 DMMA accumulators (128 FP64 registers).  Operands stream through a 4-stage cp.async
// (LDGSTS) pipeline in slices of BK = 16; shared-memory rows are padded by 4 doubles so that the
// 64-bit fragment loads of a half-warp hit 16 distinct bank pairs (conflict-free) for both the
// "MN-major" and the "K-major" operand layouts.  The kernel is FP64-pipe bound by construction:
// per BK slice a CTA moves 32 KB from L2 for 262144 FMAs (~4096 tensor-pipe cycles per SM).
//
// Replaces, in the reference, every Breeze `*` / `inv` / `\` call on blocks:
//   CORE/algebra/bcholesky.scala:111-121 (L21 = A21 inv(L11)^T ; A22 - L21 L21^T),
//   CORE/algebra/PartitionedMatrixSolvers.scala:120-169 (multi-RHS forward solve),
//   CORE/algebra/PartitionedMatrixOps.scala:280-301 (block GEMM).
#include "gemm.cuh"

namespace dgp {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, NTHREADS = 256;
constexpr int LDS_MN = BM + 4;             // doubles per k-row of an MN-major stage   (132 = 4 mod 16)
constexpr int LDS_K = BK + 4;              // doubles per mn-row of a K-major stage    ( 20 = 4 mod 16)
constexpr int STAGE_MN = BK * LDS_MN;      // 2112 doubles
constexpr int STAGE_K = BM * LDS_K;        // 2560 doubles

__device__ __forceinline__ void cp_async16(double *smem, const double *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Copy one 128 x 16 operand slice into a stage.  g points at the slice origin.
template <bool KMAJOR>
__device__ __forceinline__ void load_slice(double *s, const double *g, int64_t ld, int tid) {
  if (!KMAJOR) {
    // element (mn, k) at g[mn + k*ld]; 16 columns of 128 contiguous doubles
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * NTHREADS;
      int k = c >> 6, mc = c & 63;
      cp_async16(s + k * LDS_MN + mc * 2, g + mc * 2 + (int64_t)k * ld);
    }
  } else {
    // element (mn, k) at g[k + mn*ld]; 128 rows of 16 contiguous doubles
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * NTHREADS;
      int mn = c >> 3, kc = c & 7;
      cp_async16(s + mn * LDS_K + kc * 2, g + kc * 2 + (int64_t)mn * ld);
    }
  }
}

template <bool AK, bool BKM>
__global__ void __launch_bounds__(NTHREADS, 1) dgemm_dmma_kernel(GemmArgs p) {
  extern __shared__ __align__(16) double smem[];
  constexpr int A_STAGE = AK ? STAGE_K : STAGE_MN;
  constexpr int B_STAGE = BKM ? STAGE_K : STAGE_MN;
  double *As = smem;
  double *Bs = smem + STAGES * A_STAGE;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps, 64 x 32 each

  // ---- tile decode -----------------------------------------------------------------------
  int64_t tlin = blockIdx.x;
  if (p.reverse) tlin = (int64_t)gridDim.x - 1 - tlin;
  int ti, tj;
  if (p.lower_only) {
    ti = (int)((sqrt(8.0 * (double)tlin + 1.0) - 1.0) * 0.5);
    while ((int64_t)ti * (ti + 1) / 2 > tlin) --ti;
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tlin) ++ti;
    tj = (int)(tlin - (int64_t)ti * (ti + 1) / 2);
  } else {
    int mt = p.M / BM;
    tj = (int)(tlin / mt);
    ti = (int)(tlin % mt);
  }
  const int m0 = ti * BM, n0 = tj * BN;

  int k_lo = 0, k_hi = p.K;
  if (p.kmode == K_GE_COL) k_lo = n0;
  else if (p.kmode == K_LT_ROWEND) k_hi = min(p.K, m0 + BM);
  else if (p.kmode == K_GE_ROW) k_lo = m0;
  const int KT = (k_hi - k_lo) / BK;

  const double *A = p.A + (int64_t)blockIdx.y * p.strideA;
  const double *B = p.B + (int64_t)blockIdx.y * p.strideB;
  double *C = p.C + (int64_t)blockIdx.y * p.strideC;

  // slice origins at k = k_lo
  const double *Ag = AK ? (A + k_lo + (int64_t)m0 * p.lda) : (A + m0 + (int64_t)k_lo * p.lda);
  const double *Bg = BKM ? (B + k_lo + (int64_t)n0 * p.ldb) : (B + n0 + (int64_t)k_lo * p.ldb);
  const int64_t a_step = AK ? (int64_t)BK : (int64_t)BK * p.lda;
  const int64_t b_step = BKM ? (int64_t)BK : (int64_t)BK * p.ldb;

  // Pull the C tile towards L2 while the main loop runs (read-modify-write epilogue).
  if (p.beta != 0.0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = tid + i * NTHREADS;  // 1024 lines of 128 B: 8 per column
      int col = c >> 3, seg = c & 7;
      const double *ptr = C + m0 + seg * 16 + (int64_t)(n0 + col) * p.ldc;
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
    }
  }

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // ---- prologue: STAGES-1 slices in flight -------------------------------------------------
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      load_slice<AK>(As + s * A_STAGE, Ag + s * a_step, p.lda, tid);
      load_slice<BKM>(Bs + s * B_STAGE, Bg + s * b_step, p.ldb, tid);
    }
    cp_async_commit();
  }

  // fragment base offsets inside a stage
  const int a_off = AK ? ((wm * 64 + g) * LDS_K + t) : (t * LDS_MN + wm * 64 + g);
  const int b_off = BKM ? ((wn * 32 + g) * LDS_K + t) : (t * LDS_MN + wn * 32 + g);
  constexpr int A_I = AK ? 8 * LDS_K : 8;          // next 8 rows of the fragment grid
  constexpr int A_KK = AK ? 4 : 4 * LDS_MN;        // next k4 step
  constexpr int B_J = BKM ? 8 * LDS_K : 8;
  constexpr int B_KK = BKM ? 4 : 4 * LDS_MN;

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) {
        int s = nk % STAGES;
        load_slice<AK>(As + s * A_STAGE, Ag + (int64_t)nk * a_step, p.lda, tid);
        load_slice<BKM>(Bs + s * B_STAGE, Bg + (int64_t)nk * b_step, p.ldb, tid);
      }
      cp_async_commit();
    }
    const double *as = As + (kt % STAGES) * A_STAGE + a_off;
    const double *bs = Bs + (kt % STAGES) * B_STAGE + b_off;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = as[kk * A_KK + i * A_I];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[kk * B_KK + j * B_J];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----------------------------------------------------------------------------
  // acc[i][j][r] is C(m0 + wm*64 + i*8 + g, n0 + wn*32 + j*8 + 2t + r).
  const double alpha = p.alpha, beta = p.beta;
  double *Cw = C + m0 + wm * 64 + g + (int64_t)(n0 + wn * 32 + 2 * t) * p.ldc;
  if (beta != 0.0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double old[8][2];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) old[i][r] = Cw[i * 8 + (int64_t)(j * 8 + r) * p.ldc];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          Cw[i * 8 + (int64_t)(j * 8 + r) * p.ldc] = alpha * acc[i][j][r] + beta * old[i][r];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) Cw[i * 8 + (int64_t)(j * 8 + r) * p.ldc] = alpha * acc[i][j][r];
  }
}

template <bool AK, bool BKM>
constexpr size_t smem_bytes() {
  return (size_t)STAGES * ((AK ? STAGE_K : STAGE_MN) + (BKM ? STAGE_K : STAGE_MN)) * sizeof(double);
}

template <bool AK, bool BKM>
cudaError_t opt_in() {
  return cudaFuncSetAttribute(dgemm_dmma_kernel<AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem_bytes<AK, BKM>());
}

template <bool AK, bool BKM>
void launch(cudaStream_t stream, const GemmArgs &a, dim3 grid) {
  dgemm_dmma_kernel<AK, BKM><<<grid, NTHREADS, smem_bytes<AK, BKM>(), stream>>>(a);
}

}  // namespace

int32_t gemm_init(Ctx *ctx) {
  DGP_CUDA_OK(ctx, (opt_in<false, false>()));
  DGP_CUDA_OK(ctx, (opt_in<false, true>()));
  DGP_CUDA_OK(ctx, (opt_in<true, false>()));
  DGP_CUDA_OK(ctx, (opt_in<true, true>()));
  return DGP_OK;
}

int32_t gemm(Ctx *ctx, cudaStream_t stream, const GemmArgs &a) {
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return DGP_OK;
  if (a.M % BM || a.N % BN || a.K % BK || a.K < 0)
    return ctx->fail(DGP_ERR_SHAPE, "gemm: M=%d N=%d K=%d must be multiples of %d/%d/%d", a.M, a.N, a.K, BM, BN, BK);
  if ((a.lda & 1) || (a.ldb & 1) || ((uintptr_t)a.A & 15) || ((uintptr_t)a.B & 15))
    return ctx->fail(DGP_ERR_SHAPE, "gemm: operands must be 16-byte aligned with even leading dimensions");
  if (a.lower_only && a.M != a.N) return ctx->fail(DGP_ERR_SHAPE, "gemm: lower_only needs a square C");
  if (a.kmode != K_FULL && (a.K % BM)) return ctx->fail(DGP_ERR_SHAPE, "gemm: triangular k-range needs K %% 128 == 0");
  int64_t mt = a.M / BM, nt = a.N / BN;
  int64_t tiles = a.lower_only ? mt * (mt + 1) / 2 : mt * nt;
  if (tiles > 0x7fffffffLL || a.batch > 65535) return ctx->fail(DGP_ERR_SHAPE, "gemm: grid too large");
  dim3 grid((unsigned)tiles, (unsigned)a.batch, 1);
  if (!a.a_kmajor && !a.b_kmajor) launch<false, false>(stream, a, grid);
  else if (!a.a_kmajor && a.b_kmajor) launch<false, true>(stream, a, grid);
  else if (a.a_kmajor && !a.b_kmajor) launch<true, false>(stream, a, grid);
  else launch<true, true>(stream, a, grid);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

}  // namespace dgp
