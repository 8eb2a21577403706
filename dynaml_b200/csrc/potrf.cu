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
constexpr int LEAF_MAIN = 256;   // elimination threads (8 warps)
constexpr int LEAF_THREADS = LEAF_MAIN + 32;  // + one pivot warp
constexpr size_t LEAF_SMEM = 0;

__device__ __forceinline__ void leaf_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(LEAF_THREADS) : "memory"); }

// Cholesky + triangular inverse of one 128x128 diagonal block, register resident.
//
// The 256 elimination threads form a 16x16 grid (tr, tc); thread (tr, tc) owns the elements (r, c) with
// r = 16 i + tr, c = 16 q + tc (2D block-cyclic), i.e. an 8x8 local block of which only the local
// lower part i >= q is live (36 FP64 values for L, 36 for W = L^-1).  Each elimination step
// broadcasts one column of L and one row of W through double-buffered 128-entry lines in shared
// memory and costs a single barrier; all register indices are compile-time after unrolling
// the 16-wide column blocks, so nothing spills.
//
// The 128 reciprocal square roots form the serial chain of the factorisation; a ninth warp runs it one column
// AHEAD of the elimination: with column k the owners also publish a(k+1,k+1) as it stands before that column's
// update, and the pivot warp forms l = a(k+1,k) / l_kk, d = a(k+1,k+1) - l^2 and 1 / sqrt(d) -- the same three
// operations, in the same order and precision, that the elimination threads apply to that element -- while they
// are busy with the rank-1 updates of step k.  No elimination thread executes an rsqrt or waits for one.
//
// blockIdx.x > 0: copy CTAs.  The panel below the diagonal block is multiplied by inv(L)^T right after this
// kernel, by a GEMM whose CTAs share input rows and therefore cannot run in place; CTA b copies rows
// [128 (b-1), 128 b) of the panel (panel_src, leading dimension lda) to the scratch operand of that product
// (leading dimension mrows) underneath the factorisation.
// (nine warps put three on one SM sub-partition, whose 16 K registers then allow 168 per thread: ptxas spills ~28 values)
__global__ void __launch_bounds__(LEAF_THREADS, 1)
leaf_potrf_inv_kernel(double *A, int64_t lda, double *Inv, int *info, int pivot_base, const double *panel_src,
                      double *scratch, int64_t mrows) {
  __shared__ double colbuf[2][LB];
  __shared__ double rowbuf[2][LB];
  __shared__ double dnext[2];  // a(k+1,k+1) at the start of step k
  __shared__ double liv[2];    // 1 / l_kk of step k
  const int tid = threadIdx.x;

  if (blockIdx.x > 0) {
    if (tid < LEAF_MAIN) {
      const int64_t r = (int64_t)(blockIdx.x - 1) * LB + (tid & (LB - 1));
      const double *src = panel_src + r;
      double *dst = scratch + r;
#pragma unroll 8
      for (int c = tid >> 7; c < LB; c += 2) dst[(int64_t)c * mrows] = src[(int64_t)c * lda];
    }
    return;
  }

  if (tid >= LEAF_MAIN) {
    // ---- pivot warp (all lanes compute the same scalars; lane 0 publishes) ----------------------------------
    double d = A[0];
    if (!(d > 0.0)) {
      if (tid == LEAF_MAIN) atomicCAS(info, 0, pivot_base + 1);
      d = 1.0;
    }
    double li = rsqrt(d);
    if (tid == LEAF_MAIN) liv[0] = li;
    for (int k = 0; k < LB; ++k) {
      leaf_bar();  // column k, a(k+1,k+1) and (from here) liv[k & 1] are visible
      if (k + 1 < LB) {
        const double l = colbuf[k & 1][k + 1] * li;      // L(k+1, k)
        double dn = fma(-l, l, dnext[k & 1]);            // a(k+1,k+1) after the update of column k
        if (!(dn > 0.0)) {                               // also catches NaN
          if (tid == LEAF_MAIN) atomicCAS(info, 0, pivot_base + k + 2);
          dn = 1.0;
        }
        li = rsqrt(dn);
        if (tid == LEAF_MAIN) liv[(k + 1) & 1] = li;
      }
    }
    return;
  }

  const int tr = tid & 15, tc = tid >> 4;

  double a[8][8], w[8][8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      w[i][q] = 0.0;
      a[i][q] = (i >= q) ? A[(i * 16 + tr) + (int64_t)(q * 16 + tc) * lda] : 0.0;
    }

  // ---- factorisation and inverse, fused: one elimination step and one barrier per column ----
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
      {  // the next pivot, before this column's update, for the pivot warp
        const int kn = (k + 1) & 15;
        if (tr == kn && tc == kn && k + 1 < LB) dnext[k & 1] = (kk < 15) ? a[kc][kc] : a[(kc + 1) & 7][(kc + 1) & 7];
      }
      leaf_bar();
      double d = cb[k];
      if (!(d > 0.0)) d = 1.0;  // the pivot warp has flagged it and used the same substitute
      const double li = liv[k & 1];
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

// ------------------------------------------------------------------------------------------------------------------
// Tile leaf (the default): the same 128x128 Cholesky + inverse, blocked by 8 columns on the FP64 tensor pipe.
//
// The scalar leaf above issues ~200 instructions per thread and column, a tenth of them useful FMAs (ncu / SASS:
// 64 us per block, 4.1 of the 9.1 ms of an N = 8192 factorisation).  Here L and W = L^-1 live in shared memory as
// 136 lower 8x8 tiles each, stored in DMMA accumulator order, and one elimination step handles a whole tile column kb:
//   1. warp 0 factors the 8x8 diagonal tile (registers, unrolled) -> L88, inv(L88);
//   2. one thread per row scales the panel below, L(r, :) = A(r, :) inv(L88)^T, and one thread per column the
//      8 new rows of the inverse, W(blk, c) = inv(L88) R(blk, c); both land in fragment-ordered panel buffers;
//   3. all eight warps apply the rank-8 updates A(I,J) -= L(I,kb) L(J,kb)^T (kb < J <= I) and
//      R(I,J) -= L(I,kb) W(kb,J) (J <= kb < I) as two chained mma.sync.m8n8k4.f64 per tile, tiles dealt round-robin.
// 16 steps of three barriers instead of 128 steps of ~1000 clk.
// ------------------------------------------------------------------------------------------------------------------
constexpr int T8 = 8, NT8 = LB / T8, NTILES8 = NT8 * (NT8 + 1) / 2;  // 16 x 16 tiles, 136 on / below the diagonal
constexpr int TL_THREADS = 256;
constexpr int TL_OFF_A = 0;
constexpr int TL_OFF_W = TL_OFF_A + NTILES8 * 64;
constexpr int TL_OFF_PA = TL_OFF_W + NTILES8 * 64;   // [2][128][4]  -L(r, 4 kh + t): A operand of both updates
constexpr int TL_OFF_PB = TL_OFF_PA + 2 * LB * 4;    // [2][128][4]   L(c, 4 kh + t): B operand of the Cholesky update
constexpr int TL_OFF_PW = TL_OFF_PB + 2 * LB * 4;    // [2][128][4]   W(8 kb + 4 kh + t, c): B operand of the inverse update
constexpr int TL_OFF_I88 = TL_OFF_PW + 2 * LB * 4;   // [8][8] inverse of the diagonal tile's factor (upper part zero)
constexpr size_t TL_SMEM = (size_t)(TL_OFF_I88 + 64) * sizeof(double);

// Work items of the update phase, per elimination step kb: up to four horizontally adjacent tiles of one tile row
// (they share the A fragments).  Cholesky items first -- the very first one is the next diagonal tile, which warp 0
// takes before it factors that tile underneath the other warps' updates -- then the items of the inverse.
constexpr int TL_MAX_ITEMS = 1024;
struct ItemTab {
  unsigned int off[NT8 + 1];
  unsigned int item[TL_MAX_ITEMS];  // inverse flag << 24 | I << 16 | J0 << 8 | count
};
constexpr ItemTab make_items() {
  ItemTab t{};
  int n = 0;
  for (int kb = 0; kb < NT8; ++kb) {
    t.off[kb] = (unsigned int)n;
    for (int I = kb + 1; I < NT8; ++I)
      for (int J0 = kb + 1; J0 <= I; J0 += 4) {
        const int cnt = (I - J0 + 1) < 4 ? (I - J0 + 1) : 4;
        t.item[n++] = (unsigned int)((I << 16) | (J0 << 8) | cnt);
      }
    for (int I = kb + 1; I < NT8; ++I)
      for (int J0 = 0; J0 <= kb; J0 += 4) {
        const int cnt = (kb - J0 + 1) < 4 ? (kb - J0 + 1) : 4;
        t.item[n++] = (unsigned int)((1u << 24) | (I << 16) | (J0 << 8) | cnt);
      }
  }
  t.off[NT8] = (unsigned int)n;
  return t;
}
__constant__ ItemTab c_items = make_items();

struct TilePos {
  unsigned char I[NTILES8], J[NTILES8];
};
constexpr TilePos make_tilepos() {
  TilePos t{};
  int n = 0;
  for (int I = 0; I < NT8; ++I)
    for (int J = 0; J <= I; ++J) {
      t.I[n] = (unsigned char)I;
      t.J[n] = (unsigned char)J;
      ++n;
    }
  return t;
}
__constant__ TilePos c_tilepos = make_tilepos();   // tile index I (I + 1) / 2 + J -> (I, J)

__device__ __forceinline__ int tix(int I, int J) { return (I * (I + 1)) / 2 + J; }

__device__ __forceinline__ void dmma884_leaf(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 1 / sqrt(d) for a normal positive d: MUFU.RSQ64H seed (relative error 2^-22) and one Newton step with the cubic
// term, y (1 + e/2 + 3 e^2/8), e = 1 - d y^2 (error of the iteration 5 e^3/16 ~ 2^-67): accurate to the rounding of its
// four dependent operations, about two ulp.  The library rsqrt() adds a second step and a guarded call sequence for
// denormals and specials (which the pivot test has excluded); FP64 latency is what the 128-pivot chain consists of.
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(d));
  const double t = d * y;
  const double e = fma(-t, y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}

// Value of lane `src` of this lane's group of eight.  Raw shfl.sync: the calling warp is converged by construction
// (a warp-uniform branch), which the compiler cannot see -- __shfl_sync there is wrapped in WARPSYNC / ENDCOLLECTIVE
// pairs, tripling the instruction count of the factorisation's serial chain.
__device__ __forceinline__ double shfl8(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  // c = ((32 - width) << 8) | 0x1f  with width 8: segments of eight lanes
  asm volatile("shfl.sync.idx.b32 %0, %0, %2, 0x181f, 0xffffffff;\n\tshfl.sync.idx.b32 %1, %1, %2, 0x181f, 0xffffffff;\n"
               : "+r"(lo), "+r"(hi)
               : "r"(src));
  return __hiloint2double(hi, lo);
}

// Cholesky factor and inverse of the 8x8 tile D (row-major, lower part read) by one warp: lane & 7 is a row of
// the factorisation (and then a column of the inverse); the four groups of eight lanes work redundantly, group 0
// stores.  Results: D (in place, upper part zeroed), i88 and Wtile = inv(L).  Pivots <= 0 or NaN flag `info`.
__device__ __forceinline__ void factor_tile8(double *D, double *i88, double *Wtile, int *info, int pivot0, int lane) {
  const int i = lane & 7;
  double l[8], li[8];
#pragma unroll
  for (int b = 0; b < 8; b += 2) {
    const double2 x = *reinterpret_cast<const double2 *>(D + i * 8 + b);
    l[b] = x.x;
    l[b + 1] = x.y;
  }
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double d = shfl8(l[k], k);   // a_kk, held by the lane of row k
    const bool ok = d > 0.0;      // false for NaN too
    bad = (!ok && bad == 0) ? k + 1 : bad;
    d = ok ? d : 1.0;
    // 1 / sqrt(d) as in rsqrt_pos, with the Newton correction applied to the column entry directly:
    // l_ik = a_ik y (1 + e (1/2 + 3 e / 8)) -- one dependent operation fewer on the pivot chain than (a_ik * r)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(d));
    const double x0 = ((i == k) ? d : l[k]) * y;
    const double tt = d * y;
    const double e = fma(-tt, y, 1.0);
    const double cc = fma(0.375, e, 0.5);
    li[k] = fma(y * e, cc, y);
    const double lk = fma(x0 * e, cc, x0);
    l[k] = lk;
    // the next pivot first, from this lane's own value: its update must not wait for a shuffle
    if (k + 1 < 8 && i == k + 1) l[k + 1] = fma(-lk, lk, l[k + 1]);
#pragma unroll
    for (int j = k + 1; j < 8; ++j) {
      const double ljk = shfl8(lk, j);
      if (i > j || (i == j && j != k + 1)) l[j] = fma(-lk, ljk, l[j]);   // rows i < j hold nothing that is used
    }
  }
  if (bad != 0 && lane == 0) atomicCAS(info, 0, pivot0 + bad);
  if (lane < 8) {
#pragma unroll
    for (int b = 0; b < 8; b += 2) {
      double2 x;
      x.x = (b <= i) ? l[b] : 0.0;
      x.y = (b + 1 <= i) ? l[b + 1] : 0.0;
      *reinterpret_cast<double2 *>(D + i * 8 + b) = x;
    }
  }
  __syncwarp();
  // V = L^-1, one column per lane (column k = lane & 7), right-looking so that each row costs one multiply and one
  // independent FMA per remaining row:  v_r = -s_r / l_rr,  s_r' += l_r'r v_r  (r' > r);  v_k = 1 / l_kk, v_r = 0 above
  const int k = i;
  double v[8], sacc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) sacc[r] = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    v[r] = (r > k) ? -sacc[r] * li[r] : ((r == k) ? li[r] : 0.0);
#pragma unroll
    for (int q = r + 1; q < 8; ++q) sacc[q] = fma(D[q * 8 + r], v[r], sacc[q]);
  }
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      i88[r * 8 + k] = v[r];
      Wtile[r * 8 + k] = v[r];
    }
  }
}

__global__ void __launch_bounds__(TL_THREADS, 1)
leaf_tile_kernel(double *A, int64_t lda, double *Inv, int *info, int pivot_base, const double *panel_src,
                 double *scratch, int64_t mrows) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  if (blockIdx.x > 0) {  // copy CTAs: the panel below the block -> scratch operand of the panel product
    const int64_t r = (int64_t)(blockIdx.x - 1) * LB + (tid & (LB - 1));
    const double *src = panel_src + r;
    double *dst = scratch + r;
#pragma unroll 8
    for (int c = tid >> 7; c < LB; c += 2) dst[(int64_t)c * mrows] = src[(int64_t)c * lda];
    return;
  }
  double *sA = sm + TL_OFF_A, *sW = sm + TL_OFF_W, *pA = sm + TL_OFF_PA, *pB = sm + TL_OFF_PB, *pW = sm + TL_OFF_PW;
  double *i88 = sm + TL_OFF_I88;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // warp-uniform for the compiler as well
  const int g = lane >> 2, t = lane & 3;

  // ---- load the lower tiles, one tile per warp instruction pair: lane (g, t) fetches the two elements of its
  //      accumulator slot (eight consecutive rows of a column are one 64-byte segment), conflict-free 16-byte
  //      shared-memory stores; 34 independent loads in flight per lane.  Clear W. -----------------------------------
  {
    double2 v[NTILES8 / 8];
#pragma unroll
    for (int n = 0; n < NTILES8 / 8; ++n) {
      const int ti = warp + 8 * n;
      const int I = c_tilepos.I[ti], J = c_tilepos.J[ti];
      const double *src = A + (I * 8 + g) + (int64_t)(J * 8 + 2 * t) * lda;
      v[n].x = src[0];
      v[n].y = src[lda];
    }
#pragma unroll
    for (int n = 0; n < NTILES8 / 8; ++n) reinterpret_cast<double2 *>(sA + (warp + 8 * n) * 64)[lane] = v[n];
    for (int idx = tid; idx < NTILES8 * 32; idx += TL_THREADS) reinterpret_cast<double2 *>(sW)[idx] = make_double2(0.0, 0.0);
  }
  __syncthreads();
  if (warp == 0) factor_tile8(sA, i88, sW, info, pivot_base, lane);
  __syncthreads();

  for (int kb = 0; kb < NT8; ++kb) {
    // ---- panel below (threads 0..127: one row each) and the new rows of W (threads 128..255: one column) ---------
    if (tid < LB) {
      const int r = tid;
      if (r >= (kb + 1) * T8) {
        double *row = sA + tix(r >> 3, kb) * 64 + (r & 7) * 8;
        double x[8], yv[8];
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
          const double2 q = *reinterpret_cast<const double2 *>(row + b);
          x[b] = q.x;
          x[b + 1] = q.y;
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {  // L(r, a) = sum_{b <= a} A(r, b) inv88(a, b)
          double s = 0.0;
#pragma unroll
          for (int b = 0; b <= a; ++b) s = fma(x[b], i88[a * 8 + b], s);
          yv[a] = s;
        }
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
          double2 q, n;
          q.x = yv[b];  q.y = yv[b + 1];
          n.x = -yv[b];  n.y = -yv[b + 1];
          *reinterpret_cast<double2 *>(row + b) = q;
          *reinterpret_cast<double2 *>(pB + (b >> 2) * LB * 4 + r * 4 + (b & 3)) = q;
          *reinterpret_cast<double2 *>(pA + (b >> 2) * LB * 4 + r * 4 + (b & 3)) = n;
        }
      }
    } else {
      const int c = tid - LB;
      if (c < (kb + 1) * T8) {
        double yv[8];
        if (c >= kb * T8) {  // inside the block: W = inv88
#pragma unroll
          for (int a = 0; a < 8; ++a) yv[a] = i88[a * 8 + (c - kb * T8)];
        } else {
          double *col = sW + tix(kb, c >> 3) * 64 + (c & 7);
          double x[8];
#pragma unroll
          for (int b = 0; b < 8; ++b) x[b] = col[b * 8];
#pragma unroll
          for (int a = 0; a < 8; ++a) {  // W(8 kb + a, c) = sum_{b <= a} inv88(a, b) R(8 kb + b, c)
            double s = 0.0;
#pragma unroll
            for (int b = 0; b <= a; ++b) s = fma(i88[a * 8 + b], x[b], s);
            yv[a] = s;
          }
#pragma unroll
          for (int a = 0; a < 8; ++a) col[a * 8] = yv[a];
        }
#pragma unroll
        for (int b = 0; b < 8; b += 2) {
          double2 q;
          q.x = yv[b];  q.y = yv[b + 1];
          *reinterpret_cast<double2 *>(pW + (b >> 2) * LB * 4 + c * 4 + (b & 3)) = q;
        }
      }
    }
    __syncthreads();
    // ---- rank-8 updates on the tensor pipe.  Warp 0 updates the next diagonal tile and factors it at once (the
    //      128 reciprocal square roots are the serial chain of the block); warps 1..7 share the other items. --------
    {
      const int first = (int)c_items.off[kb], nitems = (int)c_items.off[kb + 1] - first;
      const int q0 = warp == 0 ? 0 : warp, qstep = warp == 0 ? nitems : 7, qend = warp == 0 ? (nitems > 0 ? 1 : 0) : nitems;
      for (int q = q0; q < qend; q += qstep) {
        const unsigned int item = c_items.item[first + q];
        const int I = (item >> 16) & 0xff, J0 = (item >> 8) & 0xff, cnt = item & 0xff;
        double *base = (item >> 24) ? sW : sA;
        const double *pb = (item >> 24) ? pW : pB;
        const double a0 = pA[(I * 8 + g) * 4 + t], a1 = pA[LB * 4 + (I * 8 + g) * 4 + t];
        double2 *ct = reinterpret_cast<double2 *>(base + tix(I, J0) * 64) + lane;   // tiles of a row are adjacent
        double2 cv[4];
        double b0[4], b1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int uu = u < cnt ? u : 0;
          cv[u] = ct[uu * 32];
          b0[u] = pb[((J0 + uu) * 8 + g) * 4 + t];
          b1[u] = pb[LB * 4 + ((J0 + uu) * 8 + g) * 4 + t];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dmma884_leaf(cv[u].x, cv[u].y, a0, b0[u]);
          dmma884_leaf(cv[u].x, cv[u].y, a1, b1[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u < cnt) ct[u * 32] = cv[u];
      }
      if (warp == 0 && kb + 1 < NT8) {
        __syncwarp();
        factor_tile8(sA + tix(kb + 1, kb + 1) * 64, i88, sW + tix(kb + 1, kb + 1) * 64, info, pivot_base + (kb + 1) * T8,
                     lane);
      }
    }
    __syncthreads();
  }

  // ---- write back L (upper part zeroed) and W, tile by tile in accumulator order ------------------------------------
#pragma unroll 4
  for (int n = 0; n < NT8 * NT8 / 8; ++n) {
    const int tp = warp + 8 * n, I = tp >> 4, J = tp & 15;
    double2 lv = make_double2(0.0, 0.0), wv = make_double2(0.0, 0.0);
    if (I >= J) {
      lv = reinterpret_cast<const double2 *>(sA + tix(I, J) * 64)[lane];
      wv = reinterpret_cast<const double2 *>(sW + tix(I, J) * 64)[lane];
      if (I == J) {  // the strictly upper part of a diagonal tile is scratch
        if (g < 2 * t) lv.x = wv.x = 0.0;
        if (g < 2 * t + 1) lv.y = wv.y = 0.0;
      }
    }
    double *da = A + (I * 8 + g) + (int64_t)(J * 8 + 2 * t) * lda;
    double *di = Inv + (I * 8 + g) + (J * 8 + 2 * t) * LB;
    da[0] = lv.x;
    da[lda] = lv.y;
    di[0] = wv.x;
    di[LB] = wv.y;
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
                            double *inv, int *info, double *scratch, cudaEvent_t rest_updated = nullptr) {
  for (int64_t cs = c0; cs < c0 + w; cs += LB) {
    if (cs == c0 + LB && rest_updated)   // columns beyond the first 128 received the previous block column's update later
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(st, rest_updated, 0));
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
    const int64_t rb = cs + LB, mrows = n - rb;
    // one launch: the leaf, and (extra CTAs) the copy of the panel below it to the scratch operand of the panel
    // product -- two CTAs of that GEMM share each 128-row strip, so it cannot run in place
    {
      double *ablk = A + cs + cs * lda;
      const double *psrc = A + rb + cs * lda;
      const int pivot = (int)cs, prio = ctx->stream_priority(st);
      const dim3 grid((unsigned)(1 + mrows / LB));
      if (ctx->leaf_variant == 1)
        launch_with_priority(leaf_tile_kernel, grid, dim3(TL_THREADS), TL_SMEM, st, prio, ablk, lda, invb, info, pivot, psrc,
                             scratch, mrows);
      else
        launch_with_priority(leaf_potrf_inv_kernel, grid, dim3(LEAF_THREADS), LEAF_SMEM, st, prio, ablk, lda, invb, info,
                             pivot, psrc, scratch, mrows);
    }
    ctx->launches++;
    DGP_CUDA_OK(ctx, cudaGetLastError());
    if (rb >= n) break;
    // panel: A[rb:, cs:cs+128] <- A[rb:, cs:cs+128] * inv(L_cs)^T
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
  DGP_CUDA_OK(ctx, cudaFuncSetAttribute(leaf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_SMEM));
  return DGP_OK;
}

// Outer block for a (remaining) matrix of size n.  Wider blocks amortise the per-tile prologue/epilogue of the
// trailing update, narrower ones shorten the leaf -> panel -> update chain that bounds small sizes; the driver asks
// again for every outer block with the size of what is left.  Measured on B200 (scripts/tune_nb2.py,
// profiles/tune_nb_r2.txt; ms at fixed nb = 256 / 512 / 768 / this rule): N = 8192 8.28 / 9.01 / 9.63 / 8.23,
// N = 16384 46.98 / 47.60 / 48.06 / 46.96, N = 24576 147.95 / 147.72 / 148.21 / 147.37, N = 32768 343.06 / 340.59 /
// 340.53 / 340.13.
int64_t potrf_outer_block(const Ctx *ctx, int64_t n) {
  return ctx->nb > 0 ? ctx->nb : (n >= 28672 ? 768 : (n >= 20480 ? 512 : 256));
}

int32_t potrf_blocked(Ctx *ctx, double *A, int64_t lda, int64_t n, double *inv, int *info) {
  const bool la = ctx->lookahead && n > 256;
  return potrf_blocked_on(ctx, ctx->stream, la ? ctx->panel : ctx->stream, A, lda, n, inv, info, "potrf_panel");
}

int32_t potrf_blocked_on(Ctx *ctx, cudaStream_t S, cudaStream_t P, double *A, int64_t lda, int64_t n, double *inv,
                         int *info, const char *scratch_name) {
  if (n % LB) return ctx->fail(DGP_ERR_SHAPE, "potrf: n=%lld must be a multiple of 128", (long long)n);
  // the outer block follows the size of what is left: the trailing matrix is a factorisation of that size in its own
  // right, wide blocks where the trailing GEMM sets the pace, 256 where the leaf chain does (profiles/chain_timeline_r2.txt)
  auto block_at = [&](int64_t c) { return potrf_outer_block(ctx, n - c); };
  const bool la = (P != S) && n > block_at(0);
  if (!la) P = S;
  size_t ev = 0;

  double *scratch = (double *)ctx->buf(scratch_name, (size_t)n * LB * sizeof(double));
  if (!scratch) return DGP_ERR_OOM;
  if (la) {  // the panel stream starts after whatever produced A on the main stream
    cudaEvent_t e = get_event(ctx, ev++);
    DGP_CUDA_OK(ctx, cudaEventRecord(e, S));
    DGP_CUDA_OK(ctx, cudaStreamWaitEvent(P, e, 0));
  }

  cudaEvent_t rest_ev = nullptr;  // set when the block column about to be factored was updated in two parts
  for (int64_t c0 = 0, w = 0; c0 < n; c0 += w) {
    const int64_t nb = block_at(c0);
    w = (n - c0 < nb) ? (n - c0) : nb;
    DGP_TRY(factor_block_column(ctx, P, A, lda, n, c0, w, inv, info, scratch, rest_ev));
    rest_ev = nullptr;
    const int64_t c1 = c0 + w;
    if (c1 >= n) break;
    if (la) {
      cudaEvent_t e = get_event(ctx, ev++);
      DGP_CUDA_OK(ctx, cudaEventRecord(e, P));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(S, e, 0));
    }
    const int64_t nb1 = block_at(c1), w1 = (n - c1 < nb1) ? (n - c1) : nb1;
    // (a) trailing update restricted to the next block column.  With look-ahead only its first 128 columns are on
    //     the critical chain (the next leaf and its panel need nothing else), so they go first and the panel stream
    //     resumes as soon as they are done; the rest of the block column follows and is awaited before its own
    //     128-wide sub-panels are touched.
    const int64_t wa = (la && w1 > LB) ? LB : w1;
    GemmArgs a;
    a.A = A + c1 + c0 * lda;  a.lda = lda;  a.a_kmajor = 0;
    a.B = A + c1 + c0 * lda;  a.ldb = lda;  a.b_kmajor = 0;
    a.C = A + c1 + c1 * lda;  a.ldc = lda;
    a.M = (int)(n - c1);  a.N = (int)wa;  a.K = (int)w;
    a.alpha = -1.0;  a.beta = 1.0;
    DGP_TRY(gemm(ctx, S, a));
    if (la) {
      cudaEvent_t e = get_event(ctx, ev++);
      DGP_CUDA_OK(ctx, cudaEventRecord(e, S));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(P, e, 0));
    }
    if (wa < w1) {
      GemmArgs a2 = a;   // rows from c1 + 128: the tiles above belong to the strictly upper part of the block column
      a2.A = A + (c1 + wa) + c0 * lda;
      a2.B = A + (c1 + wa) + c0 * lda;
      a2.C = A + (c1 + wa) + (c1 + wa) * lda;
      a2.M = (int)(n - c1 - wa);  a2.N = (int)(w1 - wa);
      DGP_TRY(gemm(ctx, S, a2));
      rest_ev = get_event(ctx, ev++);
      DGP_CUDA_OK(ctx, cudaEventRecord(rest_ev, S));
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

int32_t potrf_blocked_graph(Ctx *ctx, cudaStream_t S, cudaStream_t P, double *A, int64_t lda, int64_t n, double *inv,
                            int *info, const char *scratch_name) {
  if (!ctx->use_graphs) return potrf_blocked_on(ctx, S, P, A, lda, n, inv, info, scratch_name);
  const auto key = std::make_tuple((const void *)A, lda, n, (const void *)inv, (const void *)info, (const void *)S,
                                   (const void *)P);
  auto &entry = ctx->potrf_graphs[key];
  if (entry.first) {
    DGP_CUDA_OK(ctx, cudaGraphLaunch(entry.first, S));
    ctx->launches += (uint64_t)entry.second;
    return DGP_OK;
  }
  if (entry.second == 0) {  // first sight of this problem: eager (allocations, tile tables)
    entry.second = -1;
    return potrf_blocked_on(ctx, S, P, A, lda, n, inv, info, scratch_name);
  }
  // second call: capture.  Everything potrf_blocked_on touches exists by now, so no allocation or copy is issued.
  const uint64_t before = ctx->launches;
  cudaGraph_t graph = nullptr;
  DGP_CUDA_OK(ctx, cudaStreamBeginCapture(S, cudaStreamCaptureModeRelaxed));
  const int32_t st = potrf_blocked_on(ctx, S, P, A, lda, n, inv, info, scratch_name);
  const cudaError_t ce = cudaStreamEndCapture(S, &graph);
  if (st != DGP_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ctx->potrf_graphs.erase(key);
    ctx->use_graphs = 0;  // this driver / configuration cannot capture the sequence: stay eager
    ctx->launches = before;
    if (st != DGP_OK) return st;
    return potrf_blocked_on(ctx, S, P, A, lda, n, inv, info, scratch_name);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    ctx->potrf_graphs.erase(key);
    ctx->use_graphs = 0;
    ctx->launches = before;
    return potrf_blocked_on(ctx, S, P, A, lda, n, inv, info, scratch_name);
  }
  entry.first = exec;
  entry.second = (int)(ctx->launches - before);   // kernels inside the graph, for the launch counter
  DGP_CUDA_OK(ctx, cudaGraphLaunch(exec, S));
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

int32_t trsm_right_lower_t(Ctx *ctx, cudaStream_t S, cudaStream_t P, const double *L, int64_t ldl, const double *inv,
                           int64_t n, double *B, int64_t ldb, double *V, int64_t ldv, int64_t m, int64_t nb) {
  if (n % LB || nb % LB || m % LB) return ctx->fail(DGP_ERR_SHAPE, "trsm: n, nb and m must be multiples of 128");
  const bool la = (P != S) && n > 2 * nb;
  if (!la) P = S;
  size_t ev = 0;
  auto link = [&](cudaStream_t from, cudaStream_t to) -> int32_t {
    if (from == to) return DGP_OK;
    cudaEvent_t e = get_event(ctx, ev++);
    DGP_CUDA_OK(ctx, cudaEventRecord(e, from));
    DGP_CUDA_OK(ctx, cudaStreamWaitEvent(to, e, 0));
    return DGP_OK;
  };
  DGP_TRY(link(S, P));
  for (int64_t c0 = 0; c0 < n; c0 += nb) {
    const int64_t c1 = (c0 + nb < n) ? c0 + nb : n;
    // ---- the block's own columns, 128 at a time, on the panel stream (small dependent launches)
    for (int64_t cs = c0; cs < c1; cs += LB) {
      GemmArgs d;  // V_cs = B_cs * inv(L_cs,cs)^T  (out of place: the two CTAs of a 128-row strip read all 128 columns)
      d.A = B + cs * ldb;  d.lda = ldb;  d.a_kmajor = 0;
      d.B = inv + (cs / LB) * LB * LB;  d.ldb = LB;  d.b_kmajor = 0;   // op(B)(k,n) = inv(n,k)
      d.C = V + cs * ldv;  d.ldc = ldv;
      d.M = (int)m;  d.N = LB;  d.K = LB;
      d.alpha = 1.0;  d.beta = 0.0;
      DGP_TRY(gemm(ctx, P, d));
      const int64_t rb = cs + LB;
      if (rb >= c1) break;
      GemmArgs u;  // columns of the same block: B[:, rb:c1] -= V_cs * L[rb:c1, cs]^T
      u.A = V + cs * ldv;  u.lda = ldv;  u.a_kmajor = 0;
      u.B = L + rb + cs * ldl;  u.ldb = ldl;  u.b_kmajor = 0;
      u.C = B + rb * ldb;  u.ldc = ldb;
      u.M = (int)m;  u.N = (int)(c1 - rb);  u.K = LB;
      u.alpha = -1.0;  u.beta = 1.0;
      DGP_TRY(gemm(ctx, P, u));
    }
    if (c1 >= n) break;
    DGP_TRY(link(P, S));
    // ---- every column beyond the block: B[:, c1:] -= V[:, c0:c1] * L[c1:, c0:c1]^T.  With look-ahead the next
    //      block's columns go first and the panel stream resumes under the rest.
    const int64_t c2 = (c1 + nb < n) ? c1 + nb : n;
    const int64_t first = la ? c2 - c1 : n - c1;
    GemmArgs t;
    t.A = V + c0 * ldv;  t.lda = ldv;  t.a_kmajor = 0;
    t.B = L + c1 + c0 * ldl;  t.ldb = ldl;  t.b_kmajor = 0;
    t.C = B + c1 * ldb;  t.ldc = ldb;
    t.M = (int)m;  t.N = (int)first;  t.K = (int)(c1 - c0);
    t.alpha = -1.0;  t.beta = 1.0;
    DGP_TRY(gemm(ctx, S, t));
    DGP_TRY(link(S, P));
    if (c1 + first < n) {
      GemmArgs t2 = t;
      t2.B = L + c2 + c0 * ldl;
      t2.C = B + c2 * ldb;
      t2.N = (int)(n - c2);
      DGP_TRY(gemm(ctx, S, t2));
    }
  }
  DGP_TRY(link(P, S));
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
