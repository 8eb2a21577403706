// lean.cuh -- pieces shared by the lean Gram epilogues: the kernel-matrix assembly (assemble.cu) and the gradient
// trace pass (grad.cu).  Both compute |x - y|^2 as |x|^2 + |y|^2 - 2 x.y on the FP64 tensor path (m8n8k4 DMMA) from
// operands pre-scaled so that the accumulator is already the exponent's argument, and evaluate exp / sqrt with the
// short sequences below instead of the library routines.
#pragma once
#include <cmath>
#include "common.cuh"

namespace dgp {

struct LeanCov {
  // The operands are scaled by `scale` when they are staged in shared memory, so that the accumulator already
  // holds the exponent's argument: SE scale = sqrt(1/(2 l^2)) -> acc = r^2/(2 l^2); Matern scale = sqrt(2 nu)/l
  // -> sqrt(acc) = sqrt(2 nu) r / l.  This removes one FP64 multiply per entry.
  double scale;
  double m0, m1, m2, m3;  // Matern: polynomial in sqrt(acc), coefficients right-aligned (m3 is the constant term)
};

__device__ __forceinline__ double rsqrt_seed(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));  // MUFU.RSQ64H, ~2^-22 relative
  return y;
}

// exp(x) * scale for x <= ~0, scale folded into the table; exact zero below e^-700 like fast_exp_nonpos_v.
// Every instruction that is not FP64 arithmetic still costs FP64-pipe time on this part
// (scripts/ubench/fp64_mix.cu: the same FP64 sequence retires at 2.0 warp-instructions/clk/SM alone,
// 1.76 with the integer epilogue and 1.49 with the table load), so the integer side is pared down:
// the reduction constants come from the kernel parameters (constant-bank operands, no MOVs), the
// table address is one LOP3 + one shift-add, the exponent one LOP3 + one shift-add, and the underflow test
// is one integer minimum per four entries with a cold fix-up branch.  The table is replicated 16 times
// (entry j, copy c at (16 j + c) * 8 bytes) and lane l reads copy l & 15, so the 16 lanes of a half-warp hit
// 16 distinct bank pairs whatever their j: the random lookups are conflict free (the single-copy table cost
// 3.4 extra wavefronts per load, ncu l1tex__data_bank_conflicts, and ~10 % of the kernel time).
struct LeanConst {
  double l2e64, ln2hi64, ln2lo64, c5, c4, c3;
};

template <int V>
__device__ __forceinline__ void lean_exp_v(const double (&x)[V], const LeanConst &c, unsigned tab_s,
                                           double (&out)[V]) {
  const double SHIFT = 6755399441055744.0;
  int n[V];
  double r[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const double t = fma(x[v], c.l2e64, SHIFT);
    n[v] = __double2loint(t);
    const double nd = t - SHIFT;
    r[v] = fma(nd, -c.ln2lo64, fma(nd, -c.ln2hi64, x[v]));
  }
  int nmin = n[0];
#pragma unroll
  for (int v = 1; v < V; ++v) nmin = min(nmin, n[v]);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    double p = fma(r[v], c.c5, c.c4);
    p = fma(p, r[v], c.c3);
    p = fma(p, r[v], 0.5);
    p = fma(p, r[v], 1.0);
    p = fma(p, r[v], 1.0);
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(tab_s + (((unsigned)n[v] & 63u) << 7)));
    const double w = tj * p;
    const int hi = __double2hiint(w) + (n[v] >> 6) * 0x100000;  // exponent += n >> 6
    out[v] = __hiloint2double(hi, __double2loint(w));
  }
  if (nmin < -64634) {  // some x < -700: rare, kept out of the straight-line code
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (n[v] < -64634) out[v] = 0.0;
  }
}

template <int DP>
__host__ __device__ constexpr int lean_ldk() { return (DP % 8 == 0) ? DP + 4 : DP; }  // == 4 mod 8 doubles

struct ExpTab {
  double v[64];  // amplitude * 2^(j/64), built on the host per launch
};

inline LeanConst lean_constants() {
  LeanConst lc;
  lc.l2e64 = 92.33248261689366;         // 64 / ln 2
  lc.ln2hi64 = 0.01083042469326756;     // ln2/64 with the low 21 bits cleared (n * ln2hi64 is exact)
  lc.ln2lo64 = 2.9815858269852933e-12;
  lc.c5 = 1.0 / 120.0;  lc.c4 = 1.0 / 24.0;  lc.c3 = 1.0 / 6.0;
  return lc;
}

inline ExpTab lean_exp_table(double amplitude) {
  ExpTab T;
  for (int j = 0; j < 64; ++j) T.v[j] = amplitude * exp2((double)j / 64.0);
  return T;
}

}  // namespace dgp
