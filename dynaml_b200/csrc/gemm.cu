// gemm.cu -- FP64 GEMM on the sm_100a DMMA tensor path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4;
// tcgen05 has no f64 kind).
//
// One CTA computes a 128x64 tile of C with 8 DMMA warps; each warp owns 32x32 of it as 4x4 DMMA
// accumulators (64 FP64 registers), so two CTAs are resident per SM.  Two co-resident CTAs matter
// more than a bigger tile: while one CTA runs its read-modify-write epilogue or refills its
// pipeline, the other keeps the FP64 tensor pipe busy (ncu of the first 128x128 / one-CTA version:
// tensor pipe 73 % active at K = 256, 86 % at K = 4096; profiles/).
//
// Two kernels, bit-identical results (same order of the k-terms):
//   dgemm_dmma_v4_kernel  (default for MN-major A: NT and NN products, i.e. everything on the energy /
//                         gradient path) -- a ninth warp feeds the stages with cp.async.bulk row copies
//                         through full / empty mbarriers, 128-bit fragment loads, 16-byte epilogue;
//   dgemm_dmma_kernel     (K-major A: TN / TT products; option gemm_variant = 2 for everything) -- all
//                         eight warps copy with cp.async (LDGSTS) and meet at one __syncthreads per slice;
//                         rows padded by 4 doubles make its 64-bit fragment loads conflict-free.
//
// Tiles are launched in the order of a host-built tile table: super-rows of 12 x 128 rows swept
// column by column, so the ~300 concurrently resident CTAs share ~12 A panels and ~24 B panels
// and the long-K products (triangular inverse, L^-T L^-1) stream their operands from L2 instead
// of HBM; triangular k-ranges put the heavy tiles first.
//
// Replaces, in the reference, every Breeze `*` / `inv` / `\` call on blocks:
//   CORE/algebra/bcholesky.scala:111-121 (L21 = A21 inv(L11)^T ; A22 - L21 L21^T),
//   CORE/algebra/PartitionedMatrixSolvers.scala:120-169 (multi-RHS forward solve),
//   CORE/algebra/PartitionedMatrixOps.scala:280-301 (block GEMM).
#include <algorithm>
#include <tuple>

#include "gemm.cuh"

namespace dgp {

namespace {

constexpr int BM = 128, BN = 64, NTHREADS = 256;
constexpr int LDA_MN = BM + 4;  // 132 = 4 mod 16
constexpr int LDB_MN = BN + 4;  //  68 = 4 mod 16
constexpr int SUPER_ROWS = 12;

// BK = K-slice per pipeline stage and per __syncthreads: 16 with a 3-4 deep pipeline, or 32 with
// double buffering (half as many barriers; one 32-slice is ~8k cycles of DMMA per CTA pair, several
// times the L2/HBM latency it has to cover).
template <bool AK, bool BKM, int BK>
struct Cfg {
  static constexpr int LDS_K = BK + 4;  // 20 / 36 = 4 mod 16
  static constexpr int A_STAGE = AK ? BM * LDS_K : BK * LDA_MN;
  static constexpr int B_STAGE = BKM ? BN * LDS_K : BK * LDB_MN;
  // two CTAs per SM must fit in 227 KB
  static constexpr int STAGES = BK == 32 ? 2 : (((A_STAGE + B_STAGE) * 4 * 8 <= 112 * 1024) ? 4 : 3);
  static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
};

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

// Copy one ROWS x BK operand slice into a stage (ROWS = 128 for A, 64 for B).  g points at the
// slice origin.
template <bool KMAJOR, int ROWS, int BK, bool INC = false>
__device__ __forceinline__ void load_slice(double *s, const double *g, int64_t ld, int tid) {
  constexpr int CHUNKS = ROWS * BK / 2;  // 16-byte chunks
  constexpr int LD_MN = ROWS + 4;
  constexpr int LDS_K = BK + 4;
  // Chunk c = tid + i * NTHREADS: consecutive i advance the slow index by a compile-time step, so every address is
  // the previous one plus a loop-invariant stride (one 64-bit add per chunk instead of a multiply-add chain).
  if (!KMAJOR && !INC) {
    // element (mn, k) at g[mn + k*ld]; BK columns of ROWS contiguous doubles
#pragma unroll
    for (int i = 0; i < CHUNKS / NTHREADS; ++i) {
      int c = tid + i * NTHREADS;
      int k = c / (ROWS / 2), mc = c % (ROWS / 2);
      cp_async16(s + k * LD_MN + mc * 2, g + mc * 2 + (int64_t)k * ld);
    }
  } else if (!KMAJOR) {
    // the same with incremental addresses.  Measured per GEMM variant: it pays for the B tile
    // whenever B is MN-major and for the A tile in the NT kernel only (SYRK K = 768 31.4 -> 32.2 TFLOP/s), while the
    // NN / TN main loops are faster with independent address computations for A.
    constexpr int KSTEP = NTHREADS / (ROWS / 2);
    static_assert(NTHREADS % (ROWS / 2) == 0, "chunk rows must tile the thread count");
    const int k0 = tid / (ROWS / 2), mc = tid % (ROWS / 2);
    const double *gp = g + mc * 2 + (int64_t)k0 * ld;
    double *sp = s + k0 * LD_MN + mc * 2;
    const int64_t gstep = (int64_t)KSTEP * ld;
#pragma unroll
    for (int i = 0; i < CHUNKS / NTHREADS; ++i) {
      cp_async16(sp, gp);
      gp += gstep;
      sp += KSTEP * LD_MN;
    }
  } else {
    // element (mn, k) at g[k + mn*ld]; ROWS rows of BK contiguous doubles.  (The incremental form measured
    // slower here: NN 4096^3 33.0 -> 32.6 TFLOP/s, so the K-major copy keeps independent address computations.)
#pragma unroll
    for (int i = 0; i < CHUNKS / NTHREADS; ++i) {
      int c = tid + i * NTHREADS;
      int mn = c / (BK / 2), kc = c % (BK / 2);
      cp_async16(s + mn * LDS_K + kc * 2, g + kc * 2 + (int64_t)mn * ld);
    }
  }
}

template <bool AK, bool BKM, int BK>
__global__ void __launch_bounds__(NTHREADS, 2) dgemm_dmma_kernel(GemmArgs p, const uint32_t *__restrict__ tiles) {
  extern __shared__ __align__(16) double smem[];
  using C_ = Cfg<AK, BKM, BK>;
  constexpr int A_STAGE = C_::A_STAGE, B_STAGE = C_::B_STAGE, STAGES = C_::STAGES, LDS_K = C_::LDS_K;
  double *As = smem;
  double *Bs = smem + STAGES * A_STAGE;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps, 32 x 32 each

  const uint32_t tile = tiles[blockIdx.x];
  const int m0 = (int)(tile >> 16) * BM, n0 = (int)(tile & 0xffffu) * BN;

  if (p.bc_T > 0) {
    const int64_t grow = ((int64_t)(p.bc_row0 + m0 / p.bc_T) * p.bc_Pr + p.bc_pr) * p.bc_T + m0 % p.bc_T;
    const int64_t gcol = ((int64_t)(p.bc_col0 + n0 / p.bc_T) * p.bc_Pc + p.bc_pc) * p.bc_T + n0 % p.bc_T;
    if (gcol > grow + BM - 1) return;  // uniform across the CTA, before any barrier
  }

  int k_lo = 0, k_hi = p.K;
  if (p.kmode == K_GE_COL) k_lo = (n0 / 64) * 64;
  else if (p.kmode == K_LT_ROWEND) k_hi = min(p.K, m0 + BM);
  else if (p.kmode == K_GE_ROW) k_lo = m0;
  else if (p.kmode == K_LT_COLEND) k_hi = min(p.K, n0 + BN);
  const int KT = (k_hi - k_lo) / BK;

  const double *A = p.A + (int64_t)blockIdx.y * p.strideA;
  const double *B = p.B + (int64_t)blockIdx.y * p.strideB;
  double *C = p.C + (int64_t)blockIdx.y * p.strideC;

  const double *Ag = AK ? (A + k_lo + (int64_t)m0 * p.lda) : (A + m0 + (int64_t)k_lo * p.lda);
  const double *Bg = BKM ? (B + k_lo + (int64_t)n0 * p.ldb) : (B + n0 + (int64_t)k_lo * p.ldb);
  const int64_t a_step = AK ? (int64_t)BK : (int64_t)BK * p.lda;
  const int64_t b_step = BKM ? (int64_t)BK : (int64_t)BK * p.ldb;

  // Pull the C tile towards L2 while the main loop runs (read-modify-write epilogue).
  if (p.beta != 0.0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int c = tid + i * NTHREADS;  // 512 lines of 128 B: 8 per column
      int col = c >> 3, seg = c & 7;
      const double *ptr = C + m0 + seg * 16 + (int64_t)(n0 + col) * p.ldc;
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
    }
  }

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      load_slice<AK, BM, BK, (!AK && !BKM)>(As + s * A_STAGE, Ag + s * a_step, p.lda, tid);
      load_slice<BKM, BN, BK, true>(Bs + s * B_STAGE, Bg + s * b_step, p.ldb, tid);
    }
    cp_async_commit();
  }

  const int a_off = AK ? ((wm * 32 + g) * LDS_K + t) : (t * LDA_MN + wm * 32 + g);
  const int b_off = BKM ? ((wn * 32 + g) * LDS_K + t) : (t * LDB_MN + wn * 32 + g);
  constexpr int A_I = AK ? 8 * LDS_K : 8;
  constexpr int A_KK = AK ? 4 : 4 * LDA_MN;
  constexpr int B_J = BKM ? 8 * LDS_K : 8;
  constexpr int B_KK = BKM ? 4 : 4 * LDB_MN;

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nk = kt + STAGES - 1;
      if (nk < KT) {
        int s = nk % STAGES;
        load_slice<AK, BM, BK, (!AK && !BKM)>(As + s * A_STAGE, Ag + (int64_t)nk * a_step, p.lda, tid);
        load_slice<BKM, BN, BK, true>(Bs + s * B_STAGE, Bg + (int64_t)nk * b_step, p.ldb, tid);
      }
      cp_async_commit();
    }
    const double *as = As + (kt % STAGES) * A_STAGE + a_off;
    const double *bs = Bs + (kt % STAGES) * B_STAGE + b_off;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = as[kk * A_KK + i * A_I];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[kk * B_KK + j * B_J];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // acc[i][j][r] is C(m0 + wm*32 + i*8 + g, n0 + wn*32 + j*8 + 2t + r).
  const double alpha = p.alpha, beta = p.beta;
  double *Cw = C + m0 + wm * 32 + g + (int64_t)(n0 + wn * 32 + 2 * t) * p.ldc;
  if (beta != 0.0) {
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      double old[2][4][2];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) old[jj][i][r] = Cw[i * 8 + (int64_t)((jh * 2 + jj) * 8 + r) * p.ldc];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            Cw[i * 8 + (int64_t)((jh * 2 + jj) * 8 + r) * p.ldc] =
                alpha * acc[i][jh * 2 + jj][r] + beta * old[jj][i][r];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) Cw[i * 8 + (int64_t)(j * 8 + r) * p.ldc] = alpha * acc[i][j][r];
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-copy version of the kernel (gemm_variant 4, the default): operands moved by the bulk-copy (TMA) engine
// through an mbarrier pipeline, 128-bit fragment loads, 16-byte read-modify-write epilogue.
//
// Knock-out timings of the cp.async kernel above (SYRK, K = 3072, 34.2 TFLOP/s): without the per-slice
// __syncthreads 35.8, without the cp.async copies (12 LDGSTS + address arithmetic per thread and slice) 36.5,
// without both 36.8 = the DMMA peak; without the fragment loads 35.3.  The LSU-side copy instructions and the
// block-wide barrier cost ~8 % of the tensor pipe, and both go away here: a ninth warp issues, from one elected
// lane with warp-uniform addresses, one `cp.async.bulk` per operand row (rows are contiguous in global memory and
// padded in shared memory, so no tensor map is needed) and the eight DMMA warps only wait on a `full` mbarrier,
// multiply, and arrive on an `empty` one -- no warp ever waits for another except through data it needs.
// K-slices of 16 through four stages (NT), or of 32 through two (NN: half as many copies per flop; option
// gemm_v4_bk32_mask).
//
// Fragment mapping.  The m8n8k4 fragment of lane (g, t) is one element per 8-row block, i.e. one LDS.64 per
// fragment element in the kernel above.  Any permutation of the rows of the warp's 32x32 sub-tile is a valid
// assignment as long as the epilogue uses the same one, so here fragment i of lane g stands for row
// 16*(i>>1) + 2*g + (i&1): rows 2g and 2g+1 are adjacent in an MN-major stage and come from ONE LDS.128
// (rows padded to 132 / 68 doubles = 4 mod 16 keep each quarter-warp on 8 distinct 16-byte bank groups), and in
// the epilogue they are adjacent in the column-major C (16-byte loads / stores, 128 contiguous bytes per column
// and warp instruction).  K-major stages keep 64-bit loads; their rows are padded to 18 / 34 doubles, which with this
// row assignment puts each half-warp on 16 distinct banks.
// ---------------------------------------------------------------------------------------------
template <bool AK, bool BKM, int BK_>
struct CfgB {
  static constexpr int BK = BK_;        // 16 with four stages, or 32 with two (half as many copies per flop)
  static constexpr int LDS_K = BK + 2;  // 18 / 34 = 2 mod 16
  static constexpr int LDA = BM + 4;
  static constexpr int LDB = BN + 4;
  static constexpr int A_STAGE = AK ? BM * LDS_K : BK * LDA;
  static constexpr int B_STAGE = BKM ? BN * LDS_K : BK * LDB;
  static constexpr int STAGES = BK == 16 ? 4 : 2;
  static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double) + 64;  // + 2 x STAGES mbarriers
  static constexpr uint32_t STAGE_BYTES = (BM + BN) * BK * sizeof(double);  // bytes landed per stage
  static_assert(2 * (SMEM + 1024) <= 227 * 1024, "two CTAs per SM");
};
constexpr int NTHREADS_B = NTHREADS + 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const double *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// One 16-slice of DMMAs.  as / bs point at the lane's fragment origin inside the stage; `dep` folds every
// fragment register that was loaded (see the `empty` arrival in the kernel).
template <bool AK, bool BKM, int BK, bool FOLD>
__device__ __forceinline__ void mma_slice_v4(double (&acc)[4][4][2], const double *as, const double *bs, int &dep) {
  using C_ = CfgB<AK, BKM, BK>;
  constexpr int LDS_K = C_::LDS_K, LDA = C_::LDA, LDB = C_::LDB;
#pragma unroll
  for (int kk = 0; kk < BK / 4; ++kk) {
    double a[4], b[4];
    if constexpr (AK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = as[((i >> 1) * 16 + (i & 1)) * LDS_K + 4 * kk];
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double2 v = *reinterpret_cast<const double2 *>(as + kk * 4 * LDA + 16 * h);
        a[2 * h] = v.x;  a[2 * h + 1] = v.y;
      }
    }
    if constexpr (BKM) {
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[((j >> 1) * 16 + (j & 1)) * LDS_K + 4 * kk];
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double2 w = *reinterpret_cast<const double2 *>(bs + kk * 4 * LDB + 16 * h);
        b[2 * h] = w.x;  b[2 * h + 1] = w.y;
      }
    }
    if constexpr (FOLD) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dep ^= __double2hiint(a[i]) ^ __double2hiint(b[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// acc[i][j][r] is C(m0 + wm*32 + 16*(i>>1) + 2g + (i&1), n0 + wn*32 + 16*(j>>1) + 4t + 2r + (j&1));
// Cw points at C(m0 + wm*32 + 2g, n0 + wn*32 + 4t).
__device__ __forceinline__ void epilogue_v4(const double (&acc)[4][4][2], double *Cw, int64_t ldc, double alpha,
                                            double beta) {
  if (beta != 0.0) {
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {  // column block of 16
      double2 old[2][2][2];           // [j&1][r][h]
#pragma unroll
      for (int jl = 0; jl < 2; ++jl)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            old[jl][r][h] = *reinterpret_cast<const double2 *>(Cw + 16 * h + (int64_t)(16 * jp + 2 * r + jl) * ldc);
#pragma unroll
      for (int jl = 0; jl < 2; ++jl)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double2 o = old[jl][r][h];
            o.x = alpha * acc[2 * h][2 * jp + jl][r] + beta * o.x;
            o.y = alpha * acc[2 * h + 1][2 * jp + jl][r] + beta * o.y;
            *reinterpret_cast<double2 *>(Cw + 16 * h + (int64_t)(16 * jp + 2 * r + jl) * ldc) = o;
          }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double2 o;
          o.x = alpha * acc[2 * h][j][r];
          o.y = alpha * acc[2 * h + 1][j][r];
          *reinterpret_cast<double2 *>(Cw + 16 * h + (int64_t)(16 * (j >> 1) + 2 * r + (j & 1)) * ldc) = o;
        }
  }
}

// Stage hand-over (consumer warp -> bulk-copy engine), selectable for measurement (option gemm_handover):
//   HO_DEPFOLD   the arrival's address depends on every fragment register of the slice (see the main loop)
//   HO_PRODFENCE plain arrival; the producer executes fence.proxy.async.shared::cta between its acquire of `empty` and
//                the refill (the generic -> async proxy ordering the PTX memory model asks for, paid by the producer)
//   HO_CONSFENCE every consumer lane executes fence.proxy.async.shared::cta before lane 0 arrives
//   HO_NONE      plain arrival (reproduces the hazard; testing only)
enum Handover { HO_DEPFOLD = 0, HO_PRODFENCE = 1, HO_CONSFENCE = 2, HO_NONE = 3 };

template <bool AK, bool BKM, int BK, int HO>
__global__ void __launch_bounds__(NTHREADS_B, 2) dgemm_dmma_v4_kernel(GemmArgs p, const uint32_t *__restrict__ tiles) {
  extern __shared__ __align__(16) double smem[];
  using C_ = CfgB<AK, BKM, BK>;
  constexpr int A_STAGE = C_::A_STAGE, B_STAGE = C_::B_STAGE, LDS_K = C_::LDS_K, STAGES = C_::STAGES;
  constexpr int LDA = C_::LDA, LDB = C_::LDB;
  double *As = smem;
  double *Bs = smem + STAGES * A_STAGE;
  const uint32_t bars = smem_u32(smem + STAGES * (A_STAGE + B_STAGE));  // full[s] at bars + 8 s, empty[s] after them

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler as well

  const uint32_t tile = tiles[blockIdx.x];
  const int m0 = (int)(tile >> 16) * BM, n0 = (int)(tile & 0xffffu) * BN;

  if (p.bc_T > 0) {
    const int64_t grow = ((int64_t)(p.bc_row0 + m0 / p.bc_T) * p.bc_Pr + p.bc_pr) * p.bc_T + m0 % p.bc_T;
    const int64_t gcol = ((int64_t)(p.bc_col0 + n0 / p.bc_T) * p.bc_Pc + p.bc_pc) * p.bc_T + n0 % p.bc_T;
    if (gcol > grow + BM - 1) return;  // uniform across the CTA, before any barrier
  }

  int k_lo = 0, k_hi = p.K;
  if (p.kmode == K_GE_COL) k_lo = (n0 / 64) * 64;
  else if (p.kmode == K_LT_ROWEND) k_hi = min(p.K, m0 + BM);
  else if (p.kmode == K_GE_ROW) k_lo = m0;
  else if (p.kmode == K_LT_COLEND) k_hi = min(p.K, n0 + BN);
  const int KT = (k_hi - k_lo) / BK;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);             // full: the producer's expect_tx arrival + the bytes
      mbar_init(bars + 8 * (STAGES + s), 8);  // empty: one arrival per DMMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  double *C = p.C + (int64_t)blockIdx.y * p.strideC;

  if (warp == 8) {
    // ---- producer warp --------------------------------------------------------------------------------
    if (elect_one()) {
      // One bulk copy per operand row and slice.  MN-major operands have BK rows (one per k) of BM / BN doubles,
      // K-major ones BM / BN rows of BK doubles; every address is the previous one plus a uniform stride.
      constexpr int A_ROWS = AK ? BM : BK, B_ROWS = BKM ? BN : BK;
      constexpr uint32_t A_BYTES = (AK ? BK : BM) * 8, B_BYTES = (BKM ? BK : BN) * 8;
      constexpr uint32_t A_LD = (AK ? LDS_K : LDA) * 8, B_LD = (BKM ? LDS_K : LDB) * 8;  // bytes
      const double *A = p.A + (int64_t)blockIdx.y * p.strideA;
      const double *B = p.B + (int64_t)blockIdx.y * p.strideB;
      const double *ag = AK ? (A + k_lo + (int64_t)m0 * p.lda) : (A + m0 + (int64_t)k_lo * p.lda);
      const double *bg = BKM ? (B + k_lo + (int64_t)n0 * p.ldb) : (B + n0 + (int64_t)k_lo * p.ldb);
      // after the rows of one slice: MN-major advances by BK columns (the row loop already did), K-major by BK
      // doubles from the slice origin
      const int64_t a_back = AK ? (int64_t)BK - (int64_t)A_ROWS * p.lda : 0;
      const int64_t b_back = BKM ? (int64_t)BK - (int64_t)B_ROWS * p.ldb : 0;
      const uint32_t As_u = smem_u32(As), Bs_u = smem_u32(Bs);
      int s = 0;
      uint32_t parity = 1;  // of the `empty` phase a refill waits for: phase -1 counts as complete
      for (int i = 0; i < KT; ++i) {
        const uint32_t full = bars + 8 * s;
        if (i >= STAGES) {
          mbar_wait(bars + 8 * (STAGES + s), parity);
          if constexpr (HO == HO_PRODFENCE) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        mbar_expect_tx(full, C_::STAGE_BYTES);
        uint32_t dst = As_u + (uint32_t)(s * A_STAGE) * 8;
#pragma unroll 8
        for (int r = 0; r < A_ROWS; ++r) {
          bulk_copy(dst, ag, A_BYTES, full);
          dst += A_LD;
          ag += p.lda;
        }
        ag += a_back;
        dst = Bs_u + (uint32_t)(s * B_STAGE) * 8;
#pragma unroll 8
        for (int r = 0; r < B_ROWS; ++r) {
          bulk_copy(dst, bg, B_BYTES, full);
          dst += B_LD;
          bg += p.ldb;
        }
        bg += b_back;
        if (++s == STAGES) { s = 0; parity ^= 1; }
      }
    }
    return;
  }

  // ---- DMMA warps ---------------------------------------------------------------------------------------
  // Pull the C tile towards L2 for the read-modify-write epilogue; these warps have nothing else to do until the
  // first stage lands, and the producer gets to its first copies sooner.
  if (p.beta != 0.0 && !p.no_prefetch) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int c = tid + i * NTHREADS;  // 512 lines of 128 B: 8 per column
      int col = c >> 3, seg = c & 7;
      const double *ptr = C + m0 + seg * 16 + (int64_t)(n0 + col) * p.ldc;
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
    }
  }
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int a_off = AK ? ((wm * 32 + 2 * g) * LDS_K + t) : (t * LDA + wm * 32 + 2 * g);
  const int b_off = BKM ? ((wn * 32 + 2 * g) * LDS_K + t) : (t * LDB + wn * 32 + 2 * g);

  int s = 0;
  uint32_t phase = 0;
  const int zero = p.zero;
  for (int kt = 0; kt < KT; ++kt) {
    mbar_wait(bars + 8 * s, phase);
    int dep = 0;
    mma_slice_v4<AK, BKM, BK, HO == HO_DEPFOLD>(acc, As + s * A_STAGE + a_off, Bs + s * B_STAGE + b_off, dep);
    __syncwarp();
    if constexpr (HO == HO_CONSFENCE) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    // An `empty` arrival hands the stage back to the bulk-copy engine, i.e. to the ASYNC proxy, while the fragment loads
    // above are GENERIC-proxy reads.  mbarrier release / acquire orders them only against the producer thread's own
    // generic accesses; the refill it then issues is an async-proxy write and needs a proxy fence in between.  Without
    // one the refill lands before queued fragment loads have read the stage (HO_NONE: every N = 8192 factorisation
    // wrong, scripts/handover_modes.py).  Default HO_PRODFENCE: the producer executes fence.proxy.async.shared::cta
    // after acquiring `empty` -- the documented ordering, and free for the DMMA warps.  HO_DEPFOLD (round 1) made
    // the arrival's address depend on every fragment register instead (16 LOP3 per slice, 1.8 % of the Cholesky).
    if constexpr (HO == HO_DEPFOLD) {
      if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s) + (uint32_t)(dep & zero));
    } else {
      if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s));
    }
    if (++s == STAGES) { s = 0; phase ^= 1; }
  }

  double *Cw = C + m0 + wm * 32 + 2 * g + (int64_t)(n0 + wn * 32 + 4 * t) * p.ldc;
  epilogue_v4(acc, Cw, p.ldc, p.alpha, p.beta);
}

template <bool AK, bool BKM, int HO>
cudaError_t opt_in_v4() {
  cudaError_t e = cudaFuncSetAttribute(dgemm_dmma_v4_kernel<AK, BKM, 16, HO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)CfgB<AK, BKM, 16>::SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(dgemm_dmma_v4_kernel<AK, BKM, 32, HO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)CfgB<AK, BKM, 32>::SMEM);
}

template <bool AK, bool BKM>
cudaError_t opt_in() {
  cudaError_t e = cudaFuncSetAttribute(dgemm_dmma_kernel<AK, BKM, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg<AK, BKM, 16>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(dgemm_dmma_kernel<AK, BKM, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)Cfg<AK, BKM, 32>::SMEM);
  if (e != cudaSuccess) return e;
  return opt_in_v4<AK, BKM, HO_PRODFENCE>();
}

template <bool AK, bool BKM>
void launch(cudaStream_t stream, const GemmArgs &a, dim3 grid, const uint32_t *tiles, int bk, int variant, int prio) {
  if (variant == 4) {
    // the alternative hand-over modes exist for the MN-major-A layouts the path uses (NT, NN); others use the default
    auto go = [&](auto k16, auto k32) {
      if (bk == 32) launch_with_priority(k32, grid, dim3(NTHREADS_B), CfgB<AK, BKM, 32>::SMEM, stream, prio, a, tiles);
      else launch_with_priority(k16, grid, dim3(NTHREADS_B), CfgB<AK, BKM, 16>::SMEM, stream, prio, a, tiles);
    };
    if constexpr (!AK) {
      if (a.handover == HO_DEPFOLD) return go(dgemm_dmma_v4_kernel<AK, BKM, 16, HO_DEPFOLD>, dgemm_dmma_v4_kernel<AK, BKM, 32, HO_DEPFOLD>);
      if (a.handover == HO_CONSFENCE) return go(dgemm_dmma_v4_kernel<AK, BKM, 16, HO_CONSFENCE>, dgemm_dmma_v4_kernel<AK, BKM, 32, HO_CONSFENCE>);
      if (a.handover == HO_NONE) return go(dgemm_dmma_v4_kernel<AK, BKM, 16, HO_NONE>, dgemm_dmma_v4_kernel<AK, BKM, 32, HO_NONE>);
    }
    go(dgemm_dmma_v4_kernel<AK, BKM, 16, HO_PRODFENCE>, dgemm_dmma_v4_kernel<AK, BKM, 32, HO_PRODFENCE>);
  } else if (bk == 32) dgemm_dmma_kernel<AK, BKM, 32><<<grid, NTHREADS, Cfg<AK, BKM, 32>::SMEM, stream>>>(a, tiles);
  else dgemm_dmma_kernel<AK, BKM, 16><<<grid, NTHREADS, Cfg<AK, BKM, 16>::SMEM, stream>>>(a, tiles);
}

// ---- tile tables -----------------------------------------------------------------------------
// Entry = (row tile << 16) | column tile, row tiles of 128, column tiles of 64.  Super-rows of
// SUPER_ROWS row tiles are swept column by column; `bottom_first` reverses the super-row order
// (used when the k-range, hence the work, grows with the row index).
void build_tiles(int mt, int nt, bool lower, bool bottom_first, std::vector<uint32_t> &out) {
  out.clear();
  const int nsuper = (mt + SUPER_ROWS - 1) / SUPER_ROWS;
  for (int s0 = 0; s0 < nsuper; ++s0) {
    const int s = bottom_first ? nsuper - 1 - s0 : s0;
    const int r_lo = s * SUPER_ROWS, r_hi = std::min(mt, r_lo + SUPER_ROWS);
    auto ncols = [&](int ti) { return lower ? std::min(nt, 2 * ti + 2) : nt; };
    const int cmax = ncols(r_hi - 1);
    for (int tj = 0; tj < cmax; ++tj)
      for (int ti = r_lo; ti < r_hi; ++ti)
        if (tj < ncols(ti)) out.push_back(((uint32_t)ti << 16) | (uint32_t)tj);
  }
}

}  // namespace

struct TileTable {
  uint32_t *dev = nullptr;
  size_t count = 0;
};

void gemm_release(Ctx *ctx) {
  ctx->drop_graphs();  // captured GEMM launches point at these tables
  for (auto &kv : ctx->tile_tables) cudaFree(kv.second.first);
  ctx->tile_tables.clear();
}

static int32_t get_tiles(Ctx *ctx, int mt, int nt, bool lower, bool bottom_first, TileTable *out) {
  auto key = std::make_tuple(mt, nt, (int)lower, (int)bottom_first);
  auto it = ctx->tile_tables.find(key);
  if (it == ctx->tile_tables.end()) {
    std::vector<uint32_t> host;
    build_tiles(mt, nt, lower, bottom_first, host);
    uint32_t *dev = nullptr;
    DGP_CUDA_OK(ctx, cudaMalloc(&dev, std::max<size_t>(1, host.size()) * sizeof(uint32_t)));
    // synchronous copy: the table may be consumed from either stream right away
    DGP_CUDA_OK(ctx, cudaMemcpy(dev, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    it = ctx->tile_tables.emplace(key, std::make_pair(dev, host.size())).first;
  }
  out->dev = it->second.first;
  out->count = it->second.second;
  return DGP_OK;
}

int32_t gemm_prepare(Ctx *ctx, int M, int N, int lower_only, int reverse) {
  if (M <= 0 || N <= 0) return DGP_OK;
  TileTable tt;
  return get_tiles(ctx, M / BM, N / BN, lower_only != 0, reverse != 0, &tt);
}

int32_t gemm_init(Ctx *ctx) {
  DGP_CUDA_OK(ctx, (opt_in<false, false>()));
  DGP_CUDA_OK(ctx, (opt_in<false, true>()));
  DGP_CUDA_OK(ctx, (opt_in<true, false>()));
  DGP_CUDA_OK(ctx, (opt_in<true, true>()));
  DGP_CUDA_OK(ctx, (opt_in_v4<false, false, HO_DEPFOLD>()));
  DGP_CUDA_OK(ctx, (opt_in_v4<false, true, HO_DEPFOLD>()));
  DGP_CUDA_OK(ctx, (opt_in_v4<false, false, HO_CONSFENCE>()));
  DGP_CUDA_OK(ctx, (opt_in_v4<false, true, HO_CONSFENCE>()));
  DGP_CUDA_OK(ctx, (opt_in_v4<false, false, HO_NONE>()));
  DGP_CUDA_OK(ctx, (opt_in_v4<false, true, HO_NONE>()));
  return DGP_OK;
}

int32_t gemm(Ctx *ctx, cudaStream_t stream, const GemmArgs &a_in) {
  GemmArgs a = a_in;
  a.no_prefetch = ctx->gemm_no_prefetch;
  a.handover = ctx->gemm_handover;
  if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return DGP_OK;
  if (a.M % BM || a.N % BN || a.K % 16 || a.K < 0)
    return ctx->fail(DGP_ERR_SHAPE, "gemm: M=%d N=%d K=%d must be multiples of %d/%d/16", a.M, a.N, a.K, BM, BN);
  // 32-wide slices need every k-range to be a multiple of 32 (triangular ranges are multiples of 64)
  int bk = (ctx->gemm_bk == 32 && a.K % 32 == 0) ? 32 : 16;
  if ((a.lda & 1) || (a.ldb & 1) || ((uintptr_t)a.A & 15) || ((uintptr_t)a.B & 15))
    return ctx->fail(DGP_ERR_SHAPE, "gemm: operands must be 16-byte aligned with even leading dimensions");
  if (a.lower_only && a.M != a.N) return ctx->fail(DGP_ERR_SHAPE, "gemm: lower_only needs a square C");
  if (a.kmode != K_FULL && (a.K % BM)) return ctx->fail(DGP_ERR_SHAPE, "gemm: triangular k-range needs K %% 128 == 0");
  const int mt = a.M / BM, nt = a.N / BN;
  if (mt > 65535 || nt > 65535 || a.batch > 65535) return ctx->fail(DGP_ERR_SHAPE, "gemm: grid too large");
  TileTable tt;
  DGP_TRY(get_tiles(ctx, mt, nt, a.lower_only != 0, a.reverse != 0, &tt));
  if (tt.count == 0) return DGP_OK;
  dim3 grid((unsigned)tt.count, (unsigned)a.batch, 1);
  // version 3 (128-bit fragment loads / 16-byte epilogue) needs a 16-byte aligned C with an even leading dimension
  // The bulk-copy kernel needs a 16-byte aligned C with an even leading dimension (16-byte epilogue); gemm_v4_mask
  // selects the operand layouts it is used for (bit 2*a_kmajor + b_kmajor), the cp.async kernel takes the rest.
  const bool v4_layout = (ctx->gemm_v4_mask >> (2 * (a.a_kmajor ? 1 : 0) + (a.b_kmajor ? 1 : 0))) & 1;
  const int variant =
      (ctx->gemm_variant == 4 && v4_layout && !(a.ldc & 1) && !((uintptr_t)a.C & 15) && !(a.strideC & 1)) ? 4 : 2;
  const int layout = 2 * (a.a_kmajor ? 1 : 0) + (a.b_kmajor ? 1 : 0);
  if (variant == 4) bk = (((ctx->gemm_v4_bk32_mask >> layout) & 1) && a.K % 32 == 0) ? 32 : 16;
  if (!a.a_kmajor && !a.b_kmajor) launch<false, false>(stream, a, grid, tt.dev, bk, variant, ctx->stream_priority(stream));
  else if (!a.a_kmajor && a.b_kmajor) launch<false, true>(stream, a, grid, tt.dev, bk, variant, ctx->stream_priority(stream));
  else if (a.a_kmajor && !a.b_kmajor) launch<true, false>(stream, a, grid, tt.dev, bk, variant, ctx->stream_priority(stream));
  else launch<true, true>(stream, a, grid, tt.dev, bk, variant, ctx->stream_priority(stream));
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  return DGP_OK;
}

}  // namespace dgp
