// dist.cu -- 2D block-cyclic distributed Cholesky + NLL across the GPUs of one NVSwitch box
// (BASELINE config C4: N = 131072, beyond what the dense single-GPU path is meant for).
//
// One process per GPU.  The N x N kernel matrix is cut into T x T tiles; global tile (I, J) with
// I >= J lives on rank (I mod Pr, J mod Pc) of a Pr x Pc grid, as local tile (I / Pr, J / Pc) of a
// column-major local matrix.  Every rank assembles its own tiles from the replicated inputs (no
// K traffic).  Right-looking factorisation, per tile column k:
//   1. owner of (k,k): Cholesky of the diagonal tile and W_kk = L_kk^-1            (panel stream)
//   2. broadcast W_kk (+ status word) to every rank                                (NCCL)
//   3. every rank: log-det += -sum log diag(W_kk);  z_k = W_kk r_k                 (replicated)
//   4. ranks of process column k mod Pc: panel L_Ik = A_Ik W_kk^T for their tiles I > k
//   5. broadcast the panel, one chunk per process row (owner-major layout)         (NCCL)
//   6. every rank: r_I -= L_Ik z_k for all I > k (replicated forward substitution), gather the
//      tiles of its own column indices
//   7. trailing update of the local tiles A_IJ -= L_Ik L_Jk^T, I >= J > k          (main stream)
// Step 7 is a single DMMA GEMM over the local trailing matrix with a block-cyclic mask for the
// global diagonal; it runs on the main stream while steps 1-6 of column k+1 run on the
// high-priority panel stream (look-ahead: the next tile column is updated first; panel buffers are
// double-buffered).  Because z = L^-1 r and log|L| are accumulated identically on every rank from
// broadcast data, the energy 0.5 (z'z + sum log L_ii + n log 2 pi) needs no reduction and no
// backward solve.
//
// With world == 1 and Pr*Pc > 1 the same code runs all Pr*Pc ranks of the grid on ONE GPU
// ("emulated grid": broadcasts become device copies, one stream, no look-ahead) so that the
// block-cyclic index logic is parity-tested without a multi-GPU box.
//
// The reference has no counterpart: its only distributed backend is Spark RDD algebra that the
// GP models never call; the value computed is AbstractGPRegressionModel.energy
// (CORE/models/gp/AbstractGPRegressionModel.scala:516-535).
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <time.h>

#include "assemble.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "potrf.cuh"
#include "solve.cuh"

namespace dgp {

int32_t scaled_points_pub(Ctx *ctx, cudaStream_t st, const Spec &sp, const double *X, int64_t np, int dp, int d,
                          const char *bufname, const double **out);  // api.cu
int points_stay_distinct_pub(bool has_dup, const Spec &sp, int d);
int pick_gram_pub(const Ctx *ctx, const Spec &sp, const CovParams &cp, int dp, int d, double max_sq_norm);

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  // optional (present in every NCCL >= 2.4 / 2.18): communicator introspection, async-error polling, abort
  ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // the already-loaded copy (torch bundles one) wins over the system library by SONAME
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.handle = h;
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.CommCount = (decltype(api.CommCount))dlsym(h, "ncclCommCount");
      api.CommUserRank = (decltype(api.CommUserRank))dlsym(h, "ncclCommUserRank");
      api.CommGetAsyncError = (decltype(api.CommGetAsyncError))dlsym(h, "ncclCommGetAsyncError");
      api.CommAbort = (decltype(api.CommAbort))dlsym(h, "ncclCommAbort");
      api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.GroupStart ||
          !api.GroupEnd)
        api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}

struct DistState {
  int rank = 0, world = 1, Pr = 1, Pc = 1;
  bool emulate = false;  // world == 1 but Pr*Pc > 1: all grid ranks on this GPU
  ncclComm_t comm = nullptr;
  std::vector<cudaEvent_t> events;
};

#define DGP_NCCL_OK(ctx, api, expr)                                                                   \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess)                                                                            \
      return (ctx)->fail(DGP_ERR_NCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                         (api)->GetErrorString ? (api)->GetErrorString(_r) : "nccl error");           \
  } while (0)

void dist_destroy(Ctx *ctx) {
  DistState *d = ctx->dist;
  if (!d) return;
  if (d->comm) {
    NcclApi *api = nccl_api();
    if (api) api->CommDestroy(d->comm);
  }
  for (auto e : d->events) cudaEventDestroy(e);
  delete d;
  ctx->dist = nullptr;
}

namespace {

constexpr int MAXP = 8;  // process rows / columns of one box

// ---- block-cyclic index helpers (host) -----------------------------------------------------------
struct Grid {
  int64_t Nt;  // tiles per dimension
  int T, Pr, Pc;
  // first global tile index > k owned by process row p, and how many such tiles there are
  int64_t first_row(int64_t k, int p) const { return k + 1 + (((p - (k + 1)) % Pr) + Pr) % Pr; }
  int64_t count_row(int64_t k, int p) const {
    int64_t f = first_row(k, p);
    return f < Nt ? (Nt - 1 - f) / Pr + 1 : 0;
  }
  int64_t first_col(int64_t k, int p) const { return k + 1 + (((p - (k + 1)) % Pc) + Pc) % Pc; }
  int64_t count_col(int64_t k, int p) const {
    int64_t f = first_col(k, p);
    return f < Nt ? (Nt - 1 - f) / Pc + 1 : 0;
  }
  int64_t local_rows(int p) const { return (Nt - p + Pr - 1) / Pr; }  // tiles I in [0, Nt) with I % Pr == p
  int64_t local_cols(int p) const { return (Nt - p + Pc - 1) / Pc; }
};

struct ChunkMap {  // the panel of step k in owner-major layout, passed by value to kernels
  int64_t first[MAXP];   // first global tile of each process row's chunk
  int64_t count[MAXP];   // tiles in the chunk
  int64_t offset[MAXP];  // chunk start, in doubles; chunk p is a (count*T) x T column-major matrix
  int Pr, T;
};

// Storage of a rank's local matrix: only what lies on or below the GLOBAL diagonal is kept.  The local tile
// columns are grouped in chunks of STORE_G; chunk c is a column-major rectangle that starts at the first local tile row
// its first column needs (the rows above it are above the global diagonal for every column of the chunk), so it has
// its own leading dimension.  Sum over chunks = (1/2 + STORE_G / (2 nt)) of the full mt x nt rectangle: C4 on one
// GPU 71 GB instead of 137 GB.  A product that spans several chunks is issued chunk by chunk.
constexpr int STORE_G = 8;
struct Store {
  int T = 0, G = STORE_G;
  int64_t mt = 0, nt = 0;
  std::vector<int64_t> r0, off, ld;  // per chunk: first stored local tile row, offset in doubles, leading dimension
  int64_t doubles = 0;
  void plan(int64_t mt_, int64_t nt_, int T_, int Pr, int Pc, int pr, int pc) {
    T = T_;  mt = mt_;  nt = nt_;
    r0.clear();  off.clear();  ld.clear();
    doubles = 0;
    for (int64_t c0 = 0; c0 < nt; c0 += G) {
      const int64_t J = c0 * Pc + pc;                                   // first global tile column of the chunk
      int64_t first = J <= pr ? 0 : (J - pr + Pr - 1) / Pr;            // first local row li with li*Pr + pr >= J
      if (first > mt) first = mt;
      const int64_t cols = std::min<int64_t>(G, nt - c0);
      r0.push_back(first);
      off.push_back(doubles);
      ld.push_back((mt - first) * T);
      doubles += (mt - first) * T * cols * T;
    }
  }
  int chunk_of(int64_t lj) const { return (int)(lj / G); }
  int64_t ld_of(int64_t lj) const { return ld[chunk_of(lj)]; }
  // element (0, 0) of local tile (li, lj); li must not lie above the chunk's first stored row
  double *tile(double *base, int64_t li, int64_t lj) const {
    const int c = chunk_of(lj);
    return base + off[c] + (li - r0[c]) * T + (lj - (int64_t)c * G) * T * ld[c];
  }
};

// One grid rank's buffers.
struct RankBuf {
  int pr = 0, pc = 0;
  int64_t mt = 0, nt = 0;
  Store st;
  double *A = nullptr;
  double *Pfull[2] = {nullptr, nullptr};
  double *Pcol[2] = {nullptr, nullptr};
  double *Winv[2] = {nullptr, nullptr};
  double *r = nullptr, *z = nullptr, *scal = nullptr;  // scal: [0] log-det, [1] status, [2] z'z
  int *info = nullptr;
};

// ---- small kernels ---------------------------------------------------------------------------------
__global__ void set_status_kernel(double *slot, const int *info) { slot[0] = (double)info[0]; }

// scal[0] -= sum_i log W(i,i);  scal[1] = max(scal[1], status)     (one CTA, deterministic)
__global__ void __launch_bounds__(512) accumulate_logdet_kernel(const double *W, int T, double *scal) {
  __shared__ double sh[512];
  double s = 0.0;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s += log(W[i + (int64_t)i * T]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scal[0] -= sh[0];
    const double st = W[(int64_t)T * T];
    if (st > scal[1]) scal[1] = st;
  }
}

// y[row] (op)= sum_c M[row + c*ld] * v[c], c < T, for a block of 128 rows per CTA (256 threads: two
// column halves per row).  sub == 0: y = M v (z_k = W_kk r_k);  sub == 1: y -= M v.
__device__ __forceinline__ double row_dot(const double *M, int64_t ld, const double *v, int ncols) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    double m[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) m[u] = M[(int64_t)(c0 + u) * ld];
#pragma unroll
    for (int u = 0; u < 16; u += 4) {
      a0 = fma(m[u + 0], v[c0 + u + 0], a0);
      a1 = fma(m[u + 1], v[c0 + u + 1], a1);
      a2 = fma(m[u + 2], v[c0 + u + 2], a2);
      a3 = fma(m[u + 3], v[c0 + u + 3], a3);
    }
  }
  return (a0 + a1) + (a2 + a3);
}

// z = W r_k : grid = T/128 CTAs
__global__ void __launch_bounds__(256) tile_matvec_kernel(const double *W, int T, const double *rk, double *zk) {
  extern __shared__ double v[];  // T doubles + 128 partials
  double *part = v + T;
  for (int i = threadIdx.x; i < T; i += blockDim.x) v[i] = rk[i];
  __syncthreads();
  const int row = blockIdx.x * 128 + (threadIdx.x & 127), h = threadIdx.x >> 7;
  const int half = T / 2;
  const double acc = row_dot(W + row + (int64_t)(h * half) * T, T, v + h * half, half);
  if (h == 1) part[threadIdx.x & 127] = acc;
  __syncthreads();
  if (h == 0) zk[row] = acc + part[threadIdx.x];
}

// r_I -= P_I z_k for every tile I of every chunk: grid = (total tiles) * T/128 CTAs
__global__ void __launch_bounds__(256) panel_update_r_kernel(const double *P, ChunkMap cm, const double *zk, double *r) {
  extern __shared__ double v[];
  double *part = v + cm.T;
  const int T = cm.T;
  for (int i = threadIdx.x; i < T; i += blockDim.x) v[i] = zk[i];
  __syncthreads();
  const int per_tile = T / 128;
  int64_t tile = blockIdx.x / per_tile;
  const int sub = blockIdx.x % per_tile;
  int p = 0;
  while (tile >= cm.count[p]) {
    tile -= cm.count[p];
    ++p;
  }
  const int64_t I = cm.first[p] + (int64_t)cm.Pr * tile;
  const int64_t ld = cm.count[p] * T;
  const int lrow = sub * 128 + (threadIdx.x & 127), h = threadIdx.x >> 7;
  const int half = T / 2;
  const double *M = P + cm.offset[p] + tile * T + lrow + (int64_t)(h * half) * ld;
  const double acc = row_dot(M, ld, v + h * half, half);
  if (h == 1) part[threadIdx.x & 127] = acc;
  __syncthreads();
  if (h == 0) r[I * T + lrow] -= acc + part[threadIdx.x];
}

// Pcol(idx) <- the panel tile of global column index J = first_c + Pc*idx, for this rank's columns
__global__ void __launch_bounds__(256) gather_cols_kernel(const double *P, ChunkMap cm, double *Pcol, int64_t ncol,
                                                         int64_t first_c, int Pc) {
  const int T = cm.T;
  const int64_t idx = blockIdx.y;
  const int64_t J = first_c + (int64_t)Pc * idx;
  const int p = (int)(J % cm.Pr);
  const int64_t src_idx = (J - cm.first[p]) / cm.Pr;
  const double *src = P + cm.offset[p] + src_idx * T;
  const int64_t lds = cm.count[p] * T, ldd = ncol * T;
  double *dst = Pcol + idx * T;
  // one CTA copies 8 columns of the tile
  for (int c = blockIdx.x * 8; c < blockIdx.x * 8 + 8; ++c)
    for (int i = threadIdx.x; i < T; i += blockDim.x) dst[i + (int64_t)c * ldd] = src[i + (int64_t)c * lds];
}

// Wait for a stream while polling the communicator for asynchronous NCCL errors (a failed peer or link shows up
// there, not in the CUDA status); on error the communicator is aborted so the enqueued collectives cannot hang.
int32_t wait_polling(Ctx *ctx, DistState *d, NcclApi *api, cudaStream_t st) {
  if (!api || !d->comm || !api->CommGetAsyncError) {
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return DGP_OK;
  }
  for (;;) {
    const cudaError_t q = cudaStreamQuery(st);
    if (q == cudaSuccess) return DGP_OK;
    if (q != cudaErrorNotReady)
      return ctx->fail(DGP_ERR_CUDA, "dist: stream failed: %s", cudaGetErrorString(q));
    ncclResult_t ar = ncclSuccess;
    const ncclResult_t r = api->CommGetAsyncError(d->comm, &ar);
    if (r != ncclSuccess || (ar != ncclSuccess && ar != ncclInProgress)) {
      const ncclResult_t bad = r != ncclSuccess ? r : ar;
      if (api->CommAbort) api->CommAbort(d->comm);
      d->comm = nullptr;
      return ctx->fail(DGP_ERR_NCCL, "dist: asynchronous NCCL error: %s (communicator aborted)",
                       api->GetErrorString ? api->GetErrorString(bad) : "nccl error");
    }
    struct timespec ts = {0, 50000};
    nanosleep(&ts, nullptr);
  }
}

cudaEvent_t dist_event(DistState *d, size_t i) {
  while (d->events.size() <= i) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    d->events.push_back(e);
  }
  return d->events[i];
}

}  // namespace
}  // namespace dgp

using namespace dgp;

extern "C" {

int32_t dgp_dist_unique_id(void *id128) {
  if (!id128) return DGP_ERR_ARG;
  NcclApi *api = nccl_api();
  if (!api) return DGP_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return DGP_ERR_NCCL;
  memcpy(id128, &id, 128);
  return DGP_OK;
}

int32_t dgp_dist_init(dgp_ctx *ctx, const void *id128, int32_t rank, int32_t world, int32_t grid_rows,
                      int32_t grid_cols) {
  if (!ctx) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (world < 1 || rank < 0 || rank >= world || grid_rows < 1 || grid_cols < 1 || grid_rows > MAXP || grid_cols > MAXP)
    return ctx->fail(DGP_ERR_ARG, "dist_init: bad rank/world/grid (%d/%d, %dx%d)", rank, world, grid_rows, grid_cols);
  if (world > 1 && grid_rows * grid_cols != world)
    return ctx->fail(DGP_ERR_ARG, "dist_init: grid %dx%d does not match world size %d", grid_rows, grid_cols, world);
  dist_destroy(ctx);
  DistState *d = new DistState();
  d->rank = rank;
  d->world = world;
  d->Pr = grid_rows;
  d->Pc = grid_cols;
  d->emulate = (world == 1 && grid_rows * grid_cols > 1);
  if (world > 1) {
    NcclApi *api = nccl_api();
    if (!api || !id128) {
      delete d;
      return ctx->fail(DGP_ERR_NCCL, "dist_init: libnccl.so.2 could not be loaded");
    }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = api->CommInitRank(&d->comm, world, id, rank);
    if (r != ncclSuccess) {
      delete d;
      return ctx->fail(DGP_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(r) : "error");
    }
  }
  ctx->dist = d;
  return DGP_OK;
}

int32_t dgp_dist_finalize(dgp_ctx *ctx) {
  if (!ctx) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->panel);
  dist_destroy(ctx);
  return DGP_OK;
}

int32_t dgp_dist_info(dgp_ctx *ctx, int32_t *out, int32_t n) {
  if (!ctx || !out || n < 6) return DGP_ERR_ARG;
  DistState *d = ctx->dist;
  if (!d) return ctx->fail(DGP_ERR_ARG, "dist_info: call dgp_dist_init first");
  out[0] = 0;  out[1] = -1;  out[2] = d->Pr;  out[3] = d->Pc;  out[4] = d->emulate ? 1 : 0;  out[5] = 0;
  NcclApi *api = d->comm ? nccl_api() : nullptr;
  if (api) {
    int v = 0;
    if (api->CommCount && api->CommCount(d->comm, &v) == ncclSuccess) out[0] = v;
    if (api->CommUserRank && api->CommUserRank(d->comm, &v) == ncclSuccess) out[1] = v;
    if (api->GetVersion && api->GetVersion(&v) == ncclSuccess) out[5] = v;
  }
  return DGP_OK;
}

// Host-only description of the layout (tests / planning): out[0] tiles per dimension, out[1] local
// tile rows, out[2] local tile columns, out[3] lower tiles owned by `rank`, out[4] local matrix bytes.
int32_t dgp_dist_plan(int64_t n, int32_t tile, int32_t grid_rows, int32_t grid_cols, int32_t rank, int64_t *out) {
  if (!out || n <= 0 || tile <= 0 || tile % TILE || grid_rows < 1 || grid_cols < 1 || rank < 0 ||
      rank >= grid_rows * grid_cols)
    return DGP_ERR_ARG;
  Grid g{(n + tile - 1) / tile, tile, grid_rows, grid_cols};
  const int pr = rank / grid_cols, pc = rank % grid_cols;
  int64_t owned = 0;
  for (int64_t I = pr; I < g.Nt; I += grid_rows)
    for (int64_t J = pc; J < g.Nt && J <= I; J += grid_cols) ++owned;
  out[0] = g.Nt;
  out[1] = g.local_rows(pr);
  out[2] = g.local_cols(pc);
  out[3] = owned;
  Store st;
  st.plan(g.local_rows(pr), g.local_cols(pc), tile, grid_rows, grid_cols, pr, pc);
  out[4] = st.doubles * (int64_t)sizeof(double);
  return DGP_OK;
}

int32_t dgp_dist_energy(dgp_ctx *ctx, const dgp_data *D, const dgp_kernel_spec *spec, int32_t tile, double *e,
                        double *chol_ms) {
  if (!ctx || !D || !spec || !e) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  DistState *ds = ctx->dist;
  if (!ds) return ctx->fail(DGP_ERR_ARG, "dist_energy: call dgp_dist_init first");
  if (ds->world > 1 && !ds->comm)
    return ctx->fail(DGP_ERR_NCCL, "dist_energy: the communicator was aborted after an NCCL error; call dgp_dist_init again");
  if (tile < 256 || tile % 256 || tile > 2048)
    return ctx->fail(DGP_ERR_ARG, "dist_energy: tile must be a multiple of 256 in [256, 2048]");
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, D->d, &sp));
  NcclApi *api = ds->world > 1 ? nccl_api() : nullptr;

  const int T = tile;
  const Grid g{(D->N + T - 1) / T, T, ds->Pr, ds->Pc};
  const int64_t Ntot = g.Nt * T;
  const int nranks = ds->emulate ? ds->Pr * ds->Pc : 1;
  const bool la = !ds->emulate && ctx->lookahead;
  cudaStream_t S = ctx->stream;
  cudaStream_t P = la ? ctx->panel : ctx->stream;

  // ---- inputs padded to whole tiles -----------------------------------------------------------------
  const double *Xsrc = spec_uses_raw_points(sp) ? D->Xraw : D->X;
  const double *Xin = Xsrc;
  const double *rin = D->r;
  if (Ntot > D->Np) {
    double *Xp = (double *)ctx->buf("dist_X", (size_t)Ntot * D->dp * sizeof(double));
    double *rp = (double *)ctx->buf("dist_r0", (size_t)Ntot * sizeof(double));
    if (!Xp || !rp) return DGP_ERR_OOM;
    DGP_CUDA_OK(ctx, cudaMemsetAsync(Xp, 0, (size_t)Ntot * D->dp * sizeof(double), S));
    DGP_CUDA_OK(ctx, cudaMemsetAsync(rp, 0, (size_t)Ntot * sizeof(double), S));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(Xp, Xsrc, (size_t)D->Np * D->dp * sizeof(double), cudaMemcpyDeviceToDevice, S));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(rp, D->r, (size_t)D->Np * sizeof(double), cudaMemcpyDeviceToDevice, S));
    Xin = Xp;
    rin = rp;
  }
  const double *Xk = nullptr;
  DGP_TRY(scaled_points_pub(ctx, S, sp, Xin, Ntot, D->dp, D->d, "dist_Xs", &Xk));
  CovParams cp = make_cov_params(sp, true, D->d);

  // ---- per-rank buffers ------------------------------------------------------------------------------
  std::vector<RankBuf> ranks(nranks);
  const size_t panel_doubles = (size_t)Ntot * T;
  double *inv = (double *)ctx->buf("dist_inv", (size_t)T * TILE * sizeof(double));
  double *tt = (double *)ctx->buf("dist_tt", (size_t)T * T * sizeof(double));
  if (!inv || !tt) return DGP_ERR_OOM;
  for (int q = 0; q < nranks; ++q) {
    RankBuf &R = ranks[q];
    const int id = ds->emulate ? q : ds->rank;
    R.pr = id / ds->Pc;
    R.pc = id % ds->Pc;
    R.mt = g.local_rows(R.pr);
    R.nt = g.local_cols(R.pc);
    R.st.plan(R.mt, R.nt, T, ds->Pr, ds->Pc, R.pr, R.pc);
    char name[64];
    auto get = [&](const char *base, size_t doubles) -> double * {
      snprintf(name, sizeof name, "dist_%s_%d", base, q);
      return (double *)ctx->buf(name, (doubles ? doubles : 1) * sizeof(double));
    };
    R.A = get("A", (size_t)R.st.doubles);
    for (int b = 0; b < 2; ++b) {
      snprintf(name, sizeof name, "pf%d", b);
      R.Pfull[b] = get(name, panel_doubles);
      snprintf(name, sizeof name, "pc%d", b);
      R.Pcol[b] = get(name, (size_t)R.nt * T * T);
      snprintf(name, sizeof name, "wi%d", b);
      R.Winv[b] = get(name, (size_t)T * T + 8);
    }
    R.r = get("r", Ntot);
    R.z = get("z", Ntot);
    R.scal = get("scal", 8);
    R.info = (int *)get("info", 2);
    if (!R.A || !R.Pfull[0] || !R.Pfull[1] || !R.Pcol[0] || !R.Pcol[1] || !R.Winv[0] || !R.Winv[1] || !R.r || !R.z ||
        !R.scal || !R.info)
      return DGP_ERR_OOM;
  }

  // ---- assembly: every rank builds its own tiles ------------------------------------------------------
  for (RankBuf &R : ranks) {
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(R.r, rin, (size_t)Ntot * sizeof(double), cudaMemcpyDeviceToDevice, S));
    DGP_CUDA_OK(ctx, cudaMemsetAsync(R.z, 0, (size_t)Ntot * sizeof(double), S));
    DGP_CUDA_OK(ctx, cudaMemsetAsync(R.scal, 0, 8 * sizeof(double), S));
    if (R.mt == 0 || R.nt == 0) continue;
    for (size_t c = 0; c < R.st.r0.size(); ++c) {
      const int64_t c0 = (int64_t)c * R.st.G, cols = std::min<int64_t>(R.st.G, R.nt - c0), rows = R.mt - R.st.r0[c];
      if (rows <= 0) continue;
      AssembleArgs a;
      a.X1 = Xk;  a.X2 = Xk;  a.dp = D->dp;
      a.rows = D->N;  a.cols = D->N;
      a.rows_p = rows * T;  a.cols_p = cols * T;
      a.out = R.A + R.st.off[c];  a.ld = R.st.ld[c];
      a.symmetric = 1;
      a.bc_T = T;  a.bc_Pr = ds->Pr;  a.bc_Pc = ds->Pc;  a.bc_pr = R.pr;  a.bc_pc = R.pc;
      a.bc_row0 = (int)R.st.r0[c];  a.bc_col0 = (int)c0;
      a.use_gram = pick_gram_pub(ctx, sp, cp, D->dp, D->d, D->max_sq_norm);
      a.nodup = points_stay_distinct_pub(D->has_dup, sp, D->d);
      DGP_TRY(assemble(ctx, S, cp, a));
    }
  }
  DGP_CUDA_OK(ctx, cudaGetLastError());
  // The trailing update of step k on rank R, as the list of GEMMs it takes: the part in the next tile column first
  // (look-ahead), then the rest, cut at the chunk boundaries of the storage.  Used to build the tile tables ahead of
  // the timed pipeline and to issue the products.
  struct Piece {
    int64_t li, lj, rows, cols;  // local tile origin and extent
    bool next_col;               // the look-ahead part
  };
  auto trailing_pieces = [&](int64_t k, const RankBuf &R, std::vector<Piece> &out) {
    out.clear();
    const int64_t nrow = g.count_row(k, R.pr), ncol = g.count_col(k, R.pc);
    if (nrow == 0 || ncol == 0) return;
    const int64_t li0 = g.first_row(k, R.pr) / ds->Pr, lj0 = g.first_col(k, R.pc) / ds->Pc;
    const bool owns_next = la && g.first_col(k, R.pc) == k + 1;
    int64_t lj = lj0;
    const int64_t lje = lj0 + ncol;
    auto push = [&](int64_t a, int64_t b, bool nx) {
      const int c = R.st.chunk_of(a);
      const int64_t lis = std::max(li0, R.st.r0[c]);
      if (lis < li0 + nrow && b > a) out.push_back({lis, a, li0 + nrow - lis, b - a, nx});
    };
    if (owns_next) {
      push(lj, lj + 1, true);
      ++lj;
    }
    while (lj < lje) {
      const int64_t cend = std::min<int64_t>(lje, ((int64_t)R.st.chunk_of(lj) + 1) * R.st.G);
      push(lj, cend, false);
      lj = cend;
    }
  };
  std::vector<Piece> pieces;
  // tile tables of every product of the loop, built before the timed pipeline starts (get_tiles allocates and
  // copies synchronously on first use of a shape; there are O(Nt) distinct trailing shapes)
  for (int64_t k = 0; k + 1 < g.Nt; ++k)
    for (RankBuf &R : ranks) {
      const int64_t nrow = g.count_row(k, R.pr);
      if (R.pc == (int)(k % ds->Pc)) DGP_TRY(gemm_prepare(ctx, (int)(nrow * T), T, 0, 0));
      trailing_pieces(k, R, pieces);
      for (const Piece &pz : pieces) DGP_TRY(gemm_prepare(ctx, (int)(pz.rows * T), (int)(pz.cols * T), 0, 0));
    }
  DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
  size_t ev = 0;
  auto fork = [&](cudaStream_t from, cudaStream_t to) -> int32_t {
    if (from == to) return DGP_OK;
    cudaEvent_t evt = dist_event(ds, ev++);
    DGP_CUDA_OK(ctx, cudaEventRecord(evt, from));
    DGP_CUDA_OK(ctx, cudaStreamWaitEvent(to, evt, 0));
    return DGP_OK;
  };
  DGP_TRY(fork(S, P));

  // events of the look-ahead pipeline (real mode only)
  cudaEvent_t ev_col = la ? dist_event(ds, ev++) : nullptr;      // next tile column updated
  cudaEvent_t ev_bdone[2] = {la ? dist_event(ds, ev++) : nullptr, la ? dist_event(ds, ev++) : nullptr};
  cudaEvent_t ev_panel = la ? dist_event(ds, ev++) : nullptr;
  bool bdone_recorded[2] = {false, false};
  bool col_recorded = false;

  auto bcast = [&](auto getbuf, size_t count, int root_id) -> int32_t {
    if (count == 0) return DGP_OK;
    if (ds->emulate) {
      for (int q = 0; q < nranks; ++q)
        if (q != root_id)
          DGP_CUDA_OK(ctx, cudaMemcpyAsync(getbuf(ranks[q]), getbuf(ranks[root_id]), count * sizeof(double),
                                           cudaMemcpyDeviceToDevice, P));
    } else if (ds->world > 1) {
      double *buf = getbuf(ranks[0]);
      DGP_NCCL_OK(ctx, api, api->Broadcast(buf, buf, count, ncclDouble, root_id, ds->comm, P));
    }
    return DGP_OK;
  };

  const size_t mv_smem = ((size_t)T + 128) * sizeof(double);
  NvtxRange r_dist("dgp:dist_cholesky");
  for (int64_t k = 0; k < g.Nt; ++k) {
    const int b = (int)(k & 1);
    const int okr = (int)(k % ds->Pr), okc = (int)(k % ds->Pc);
    const int root_diag = okr * ds->Pc + okc;
    ChunkMap cm;
    cm.Pr = ds->Pr;
    cm.T = T;
    int64_t total_tiles = 0;
    {
      int64_t off = 0;
      for (int p = 0; p < MAXP; ++p) {
        cm.first[p] = p < ds->Pr ? g.first_row(k, p) : 0;
        cm.count[p] = p < ds->Pr ? g.count_row(k, p) : 0;
        cm.offset[p] = off;
        off += cm.count[p] * T * T;
        total_tiles += cm.count[p];
      }
    }
    if (la) {
      // the buffers of parity b were last read by the trailing update of step k-2; the tile column k
      // was last written by the "next column" part of step k-1
      if (bdone_recorded[b]) DGP_CUDA_OK(ctx, cudaStreamWaitEvent(P, ev_bdone[b], 0));
      if (col_recorded) DGP_CUDA_OK(ctx, cudaStreamWaitEvent(P, ev_col, 0));
      col_recorded = false;
    }
    // 1. diagonal tile: Cholesky and inverse on its owner
    for (int q = 0; q < nranks; ++q) {
      RankBuf &R = ranks[q];
      if (R.pr != okr || R.pc != okc) continue;
      double *Akk = R.st.tile(R.A, k / ds->Pr, k / ds->Pc);
      const int64_t ldk = R.st.ld_of(k / ds->Pc);
      DGP_CUDA_OK(ctx, cudaMemsetAsync(R.info, 0, sizeof(int), P));
      DGP_TRY(potrf_blocked_on(ctx, P, P, Akk, ldk, T, inv, R.info, "dist_potrf_panel"));
      DGP_CUDA_OK(ctx, cudaMemsetAsync(R.Winv[b], 0, ((size_t)T * T + 8) * sizeof(double), P));
      DGP_TRY(trtri_lower_on(ctx, P, Akk, ldk, inv, T, R.Winv[b], T, tt, T));
      set_status_kernel<<<1, 1, 0, P>>>(R.Winv[b] + (size_t)T * T, R.info);
      ctx->launches++;
    }
    // 2. W_kk to everyone
    DGP_TRY(bcast([&](RankBuf &R) { return R.Winv[b]; }, (size_t)T * T + 8, ds->emulate ? root_diag : root_diag));
    // 3. replicated log-det and z_k
    for (RankBuf &R : ranks) {
      accumulate_logdet_kernel<<<1, 512, 0, P>>>(R.Winv[b], T, R.scal);
      tile_matvec_kernel<<<T / 128, 256, mv_smem, P>>>(R.Winv[b], T, R.r + k * T, R.z + k * T);
      ctx->launches += 2;
    }
    DGP_CUDA_OK(ctx, cudaGetLastError());
    if (total_tiles == 0) break;
    // 4. panel tiles on the owning process column
    for (RankBuf &R : ranks) {
      if (R.pc != okc || cm.count[R.pr] == 0) continue;
      const int64_t li0 = cm.first[R.pr] / ds->Pr;
      GemmArgs pg;
      pg.A = R.st.tile(R.A, li0, k / ds->Pc);  pg.lda = R.st.ld_of(k / ds->Pc);  pg.a_kmajor = 0;
      pg.B = R.Winv[b];  pg.ldb = T;  pg.b_kmajor = 0;  // op(B)(k,n) = W(n,k)
      pg.C = R.Pfull[b] + cm.offset[R.pr];  pg.ldc = cm.count[R.pr] * T;
      pg.M = (int)(cm.count[R.pr] * T);  pg.N = T;  pg.K = T;
      DGP_TRY(gemm(ctx, P, pg));
    }
    // 5. the whole panel to everyone, one chunk per process row
    if (!ds->emulate && ds->world > 1) DGP_NCCL_OK(ctx, api, api->GroupStart());
    for (int p = 0; p < ds->Pr; ++p) {
      const size_t off = (size_t)cm.offset[p];
      DGP_TRY(bcast([&](RankBuf &R) { return R.Pfull[b] + off; }, (size_t)cm.count[p] * T * T, p * ds->Pc + okc));
    }
    if (!ds->emulate && ds->world > 1) DGP_NCCL_OK(ctx, api, api->GroupEnd());
    // 6. replicated forward substitution, and the tiles of this rank's own column indices
    for (RankBuf &R : ranks) {
      panel_update_r_kernel<<<(unsigned)(total_tiles * (T / 128)), 256, mv_smem, P>>>(R.Pfull[b], cm, R.z + k * T, R.r);
      ctx->launches++;
      const int64_t ncol = g.count_col(k, R.pc);
      if (ncol > 0) {
        dim3 grid((unsigned)(T / 8), (unsigned)ncol);
        gather_cols_kernel<<<grid, 256, 0, P>>>(R.Pfull[b], cm, R.Pcol[b], ncol, g.first_col(k, R.pc), ds->Pc);
        ctx->launches++;
      }
    }
    DGP_CUDA_OK(ctx, cudaGetLastError());
    if (la) {
      DGP_CUDA_OK(ctx, cudaEventRecord(ev_panel, P));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(S, ev_panel, 0));
    }
    // 7. trailing update of the local tiles: first the next tile column, then the rest chunk by chunk
    for (RankBuf &R : ranks) {
      const int64_t nrow = cm.count[R.pr], ncol = g.count_col(k, R.pc);
      if (nrow == 0 || ncol == 0) continue;
      const int64_t li0 = cm.first[R.pr] / ds->Pr, lj0 = g.first_col(k, R.pc) / ds->Pc;
      trailing_pieces(k, R, pieces);
      for (const Piece &pz : pieces) {
        GemmArgs u;
        u.A = R.Pfull[b] + cm.offset[R.pr] + (pz.li - li0) * T;  u.lda = nrow * T;  u.a_kmajor = 0;
        u.B = R.Pcol[b] + (pz.lj - lj0) * T;  u.ldb = ncol * T;  u.b_kmajor = 0;
        u.C = R.st.tile(R.A, pz.li, pz.lj);  u.ldc = R.st.ld_of(pz.lj);
        u.M = (int)(pz.rows * T);  u.N = (int)(pz.cols * T);  u.K = T;
        u.alpha = -1.0;  u.beta = 1.0;
        u.bc_T = T;  u.bc_Pr = ds->Pr;  u.bc_Pc = ds->Pc;  u.bc_pr = R.pr;  u.bc_pc = R.pc;
        u.bc_row0 = (int)pz.li;  u.bc_col0 = (int)pz.lj;
        DGP_TRY(gemm(ctx, S, u));
        if (pz.next_col) {
          // ranks outside the next process column never touch tile column k+1: their panel stream has nothing to
          // wait for and joins the next broadcasts right away
          DGP_CUDA_OK(ctx, cudaEventRecord(ev_col, S));
          col_recorded = true;
        }
      }
    }
    if (la) {
      DGP_CUDA_OK(ctx, cudaEventRecord(ev_bdone[b], S));
      bdone_recorded[b] = true;
    }
  }
  DGP_TRY(fork(P, S));
  DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
  for (RankBuf &R : ranks) DGP_TRY(dot(ctx, S, R.z, R.z, Ntot, R.scal + 2));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scal, ranks[0].scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, S));
  if (nranks > 1)
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scal + 8, ranks[nranks - 1].scal, 8 * sizeof(double),
                                     cudaMemcpyDeviceToHost, S));
  DGP_TRY(wait_polling(ctx, ds, api, S));
  DGP_TRY(wait_polling(ctx, ds, api, P));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  if (chol_ms) *chol_ms = ms;
  if (ctx->h_scal[1] != 0.0 || (nranks > 1 && ctx->h_scal[9] != 0.0)) {
    *e = INFINITY;
    return ctx->fail(DGP_ERR_NOT_PD, "distributed Cholesky failed: a pivot is not positive");
  }
  if (nranks > 1 && (ctx->h_scal[0] != ctx->h_scal[8] || ctx->h_scal[2] != ctx->h_scal[10]))
    return ctx->fail(DGP_ERR_NCCL, "emulated grid ranks disagree: logdet %.17g vs %.17g, z'z %.17g vs %.17g",
                     ctx->h_scal[0], ctx->h_scal[8], ctx->h_scal[2], ctx->h_scal[10]);
  // AbstractGPRegressionModel.scala:527-529
  *e = 0.5 * (ctx->h_scal[2] + ctx->h_scal[0] + (double)D->N * log(2.0 * M_PI));
  return DGP_OK;
}

}  // extern "C"
