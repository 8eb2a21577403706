// common.cuh -- context, workspace arena and small helpers shared by every translation unit of
// libdgp_b200.so.  Host-side plumbing only; the kernels live in the other files.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost a null-pointer test unless a profiler is attached

#include "../../include/dgp_b200.h"

namespace dgp {

constexpr int TILE = 128;  // edge of the leaf blocks / GEMM CTA tiles; every padded dimension is a multiple

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// Context: one GPU, two streams (main + high-priority panel stream for look-ahead), a grow-only
// arena of named device buffers so that repeated energy evaluations never call cudaMalloc.
// ---------------------------------------------------------------------------------------------
struct DistState;  // dist.cu

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;  // main stream (trailing updates, assembly, solves)
  cudaStream_t panel = nullptr;   // high-priority stream (leaf factorisations, panels) for look-ahead
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  std::vector<cudaEvent_t> ev_pool;  // look-ahead fork/join events
  std::vector<cudaEvent_t> ev_misc;  // fork/join of the solves running under the triangular inverse
  cudaEvent_t tev[16];               // phase timing events
  std::string err;
  std::map<std::string, std::pair<void *, size_t>> arena;
  // GEMM tile tables on the device, keyed by (row tiles, column tiles, lower, bottom_first)
  std::map<std::tuple<int, int, int, int>, std::pair<uint32_t *, size_t>> tile_tables;
  int64_t nb = 0;     // outer Cholesky block (0 = choose from n)
  int lookahead = 1;
  int gemm_handover = 1;  // how a DMMA warp hands a stage back to the bulk-copy engine (gemm.cu, Handover): producer-side proxy fence
  int gemm_variant = 4;  // 4 = bulk-copy / mbarrier kernel (dgemm_dmma_v4_kernel), 2 = cp.async kernel (dgemm_dmma_kernel)
  int gemm_v4_bk32_mask = 2; // layouts whose bulk-copy kernel uses 32-wide slices / two stages (NN: 34.6 vs 33.9 TFLOP/s)
  int gemm_v4_mask = 3;  // operand layouts (bit 2*a_kmajor + b_kmajor) that use the bulk-copy kernel: NT and NN (TN: 32.2 vs 33.9 for cp.async)
  int inverse_transposed = 1;  // gradient path: K^-1 = U U^T from U = L^-T (all NT / NN products) instead of W^T W from W = L^-1
  int leaf_variant = 1;  // 128-block leaf of the Cholesky: 1 = tile / DMMA kernel, 0 = scalar register kernel
  int trsv_chained = 1;  // triangular solves as one chained launch per direction (0: one launch per 128-block)
  int trsv_epoch = 0;    // flag value of the next chained solve
  int gemm_no_prefetch = 0;  // testing only: GEMMs skip the L2 prefetch of their C tile
  int gemm_bk = 32;   // K-slice of the GEMM pipeline: 16 (3-4 stages) or 32 (double buffered)
  int assembly = 0;   // 0: Gram/DMMA assembly when its error bound allows, 1: always exact differences, 2: force Gram
  int grad_lean = 1;       // gradient trace pass: 1 = lean Gram epilogue where the lean assembly's conditions hold, 0 = exact differences
  int assembly_store = 0;  // lean SE assembly: 0 = 16-byte st.global from registers, 1 = staged chunk + bulk (TMA) stores
  int batch_streams = 0;  // dgp_energy_batch evaluates this many candidates concurrently (1..4; 0 = by size: 4 up to N = 12288, else 3)
  cudaStream_t slot_S[4] = {nullptr, nullptr, nullptr, nullptr};  // extra (main, panel) stream pairs, created lazily
  cudaStream_t slot_P[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t slot_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int timing = 1;     // record per-phase events
  double timings[16];
  uint64_t launches = 0;  // kernels launched by this library on this context
  int *d_info = nullptr;  // device-side status word(s): [0] = first failing pivot index + 1, 0 = ok
  int *h_info = nullptr;  // pinned mirror
  double *h_scal = nullptr;  // pinned scratch for scalar read-back (64 doubles)
  cudaMemPool_t data_pool = nullptr;  // the library's own stream-ordered pool for dgp_data blocks (never trimmed at syncs)
  void *h_stage = nullptr;   // pinned, grow-only staging buffer of dgp_data_create
  size_t h_stage_bytes = 0;
  // large results leave through two pinned staging buffers on their own stream (copy_out_2d in api.cu)
  double *h_out[2] = {nullptr, nullptr};
  cudaStream_t out_stream = nullptr;
  cudaEvent_t out_ev[2] = {nullptr, nullptr};
  DistState *dist = nullptr;
  // CUDA graphs of the blocked Cholesky (potrf.cu): one executable graph per (matrix, size, streams); the
  // factorisation is ~5 launches per 128 columns, all with arguments that do not change between evaluations
  int use_graphs = 0;  // opt-in (option "graphs"): measured 3-12 % slower than live streams on one GPU, see DESIGN.md 4.2
  int prio_lo = 0, prio_hi = 0;  // stream priority range of the device (main streams: lo, panel streams: hi)
  int stream_priority(cudaStream_t st) const {
    if (st == panel) return prio_hi;
    for (int i = 0; i < 4; ++i)
      if (slot_P[i] && st == slot_P[i]) return prio_hi;
    return prio_lo;
  }
  std::map<std::tuple<const void *, int64_t, int64_t, const void *, const void *, const void *, const void *>,
           std::pair<cudaGraphExec_t, int>> potrf_graphs;   // exec (nullptr until the 2nd call), calls seen
  void drop_graphs();

  int32_t fail(int32_t code, const char *fmt, ...);
  // Named grow-only buffer.  Returns nullptr (and sets err) on OOM.
  void *buf(const char *name, size_t bytes);
  void release(const char *name);
  void release_all();
};

// Launch with the stream's priority attached as a launch attribute.  On a live stream this changes nothing (the stream
// already has that priority); inside a stream capture it is what carries the priority into the graph's kernel node --
// nodes captured from a high-priority stream otherwise run at default priority and the look-ahead panels of the
// Cholesky queue up behind the trailing update.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_with_priority(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        int priority, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributePriority;
  attr[0].val.priority = priority;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// NVTX range over the host-side enqueue of one phase: shows up in Nsight timelines next to the kernels the
// phase launched; free without a tool attached.
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
  NvtxRange &operator=(const NvtxRange &) = delete;
};

#define DGP_CUDA_OK(ctx, expr)                                                               \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return (ctx)->fail(_e == cudaErrorMemoryAllocation ? DGP_ERR_OOM : DGP_ERR_CUDA,       \
                         "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define DGP_TRY(expr)                \
  do {                               \
    int32_t _s = (expr);             \
    if (_s != DGP_OK) return _s;     \
  } while (0)

// Device-resident, padded training set.
struct Data {
  int64_t N = 0;    // real points
  int64_t Np = 0;   // padded to a multiple of the outer block
  int d = 0;        // real feature count
  int dp = 0;       // padded to a multiple of 4 (zero columns; exact for every distance used here)
  double *block = nullptr;  // one stream-ordered allocation holding X, Xraw and r
  double *X = nullptr;  // Np x dp, row-major (point i at X + i*dp), pad rows zero
  double *Xraw = nullptr;  // the same points without the centring, for the kernels that are not translation invariant
  double *r = nullptr;  // y - m(X), length Np, pad entries zero
  // The stored points are centred (column means subtracted): every kernel on this path is
  // translation invariant, and small norms keep the Gram-form distances accurate.
  std::vector<double> center;  // d column means of the original inputs
  double max_sq_norm = 0.0;    // max_i |x_i - center|^2
  bool has_dup = true;         // two training points coincide (decides where Dirac noise can land)
  // GPBasisFuncRegressionModel: Psi, column-major Np x basis_mp (zero padded), or nullptr
  double *psi = nullptr;
  int basis_m = 0, basis_mp = 0;
};

struct Composite;  // covfn.cuh

// Host copy of a kernel spec with derived constants.  For DGP_KERNEL_COMPOSITE `n_hyp` is the total
// number of base-kernel hyper-parameters (the gradient length minus one) and `comp` holds the parsed
// expression; `hyp` is unused.
struct Spec {
  int kind = 0;
  int n_hyp = 0;
  int grad_mode = 0;
  double hyp[DGP_MAX_HYP];
  double noise = 0.0;
  std::shared_ptr<const Composite> comp;
};

int32_t parse_spec(Ctx *ctx, const dgp_kernel_spec *in, int d, Spec *out);
bool spec_uses_raw_points(const Spec &sp);  // PolynomialKernel / FBMKernel somewhere in the expression

}  // namespace dgp

struct dgp_ctx : dgp::Ctx {};
struct dgp_data : dgp::Data {};
