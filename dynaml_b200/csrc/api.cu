// api.cu -- the C ABI of libdgp_b200.so (include/dgp_b200.h): context, resident training data,
// energy / gradEnergy / batched energies, fit + predict, kernel-matrix export, and the
// measurement / building-block hooks used by tests and bench.py.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <utility>

#include "assemble.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "grad.cuh"
#include "potrf.cuh"
#include "solve.cuh"

namespace dgp {

int32_t measure_peaks(Ctx *ctx, double *out, int n);  // peaks.cu
void dist_destroy(Ctx *ctx);                          // dist.cu

// ------------------------------------------------------------------------------------------------
int32_t Ctx::fail(int32_t code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err = buf;
  return code;
}

void *Ctx::buf(const char *name, size_t bytes) {
  auto it = arena.find(name);
  if (it != arena.end() && it->second.second >= bytes) return it->second.first;
  if (it != arena.end()) {
    cudaDeviceSynchronize();  // any slot's streams may still be reading the old buffer
    drop_graphs();            // captured launches hold the old address
    cudaFree(it->second.first);
    arena.erase(it);
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(DGP_ERR_OOM, "cudaMalloc(%zu bytes) for '%s' failed: %s", bytes, name, cudaGetErrorString(e));
    return nullptr;
  }
  arena[name] = {p, bytes};
  return p;
}

void Ctx::drop_graphs() {
  for (auto &kv : potrf_graphs)
    if (kv.second.first) cudaGraphExecDestroy(kv.second.first);
  potrf_graphs.clear();
}

void Ctx::release(const char *name) {
  auto it = arena.find(name);
  if (it == arena.end()) return;
  cudaDeviceSynchronize();
  drop_graphs();
  cudaFree(it->second.first);
  arena.erase(it);
}

void Ctx::release_all() {
  cudaDeviceSynchronize();
  drop_graphs();
  for (auto &kv : arena) cudaFree(kv.second.first);
  arena.clear();
}

#define DGP_BUF(var, type, ctx, name, bytes)             \
  type *var = (type *)(ctx)->buf(name, bytes);           \
  if (!var) return DGP_ERR_OOM

namespace {

int32_t parse_simple(Ctx *ctx, const dgp_kernel_spec *in, int d, Spec *out);

// DGP_KERNEL_COMPOSITE: decode the flat program (layout in dgp_b200.h) into a Composite.
int32_t parse_composite(Ctx *ctx, const dgp_kernel_spec *in, int d, Spec *out) {
  const double *p = in->hyp;
  const int n = in->n_hyp;
  auto as_int = [](double v, int lo, int hi, int *o) {
    if (!(v >= lo && v <= hi) || v != floor(v)) return false;
    *o = (int)v;
    return true;
  };
  int F = 0, T = 0;
  if (n < 2 || !as_int(p[0], 1, MAXF, &F) || !as_int(p[1], 1, MAXT, &T))
    return ctx->fail(DGP_ERR_ARG, "composite kernel: needs 1..%d base kernels and 1..%d terms", MAXF, MAXT);
  auto comp = std::make_shared<Composite>();
  CompositeDev &dev = comp->dev;
  dev.F = F;
  dev.T = T;
  const int dp = (int)round_up(d, 4);
  int pos = 2, slots = 0, total_hyp = 0, moments = 0;
  for (int s = 0; s < MAX_ARD_SLOTS; ++s) {
    dev.slot_l1[s] = 0;
    for (int c = 0; c < MAX_ARD_DIM; ++c) dev.w[s][c] = 0.0;
  }
  for (int f = 0; f < MAXF; ++f) {
    dev.metric[f] = 0;
    dev.ard_exact[f] = 0;
    dev.moment_off[f] = 0;
    dev.dim0[f] = 0;
    dev.dimn[f] = d;
    comp->factor_d[f] = d;
  }
  // ---- optional trailer: per base kernel the (start, count) slice of the input vector it sees ----------------
  {
    int q = 2;
    bool ok = true;
    for (int f = 0; f < F && ok; ++f) {
      int nf = 0;
      ok = q + 3 <= n && as_int(p[q + 2], 0, DGP_MAX_HYP, &nf);
      q += 3 + nf;
    }
    q += T * (1 + F);
    if (ok && q + 2 * F == n) {
      for (int f = 0; f < F; ++f) {
        int s0 = 0, c0 = 0;
        if (!as_int(p[q + 2 * f], 0, d - 1, &s0) || !as_int(p[q + 2 * f + 1], 0, d, &c0) || s0 + c0 > d)
          return ctx->fail(DGP_ERR_ARG, "composite kernel: bad coordinate range for base kernel %d (d = %d)", f, d);
        if (c0 == 0) c0 = d - s0;
        dev.dim0[f] = s0;
        dev.dimn[f] = c0;
        comp->factor_d[f] = c0;
      }
    }
  }
  // a weighted-distance slot: reuse an identical one, else take the next
  auto take_slot = [&](const double *w, int l1) -> int {
    for (int s = 0; s < slots; ++s) {
      bool same = dev.slot_l1[s] == l1;
      for (int c = 0; c < MAX_ARD_DIM && same; ++c) same = dev.w[s][c] == w[c];
      if (same) return s;
    }
    if (slots >= MAX_ARD_SLOTS) return -1;
    for (int c = 0; c < MAX_ARD_DIM; ++c) dev.w[slots][c] = w[c];
    dev.slot_l1[slots] = l1;
    return slots++;
  };
  for (int f = 0; f < F; ++f) {
    int kind = 0, gm = 0, nf = 0;
    if (pos + 3 > n || !as_int(p[pos], 0, DGP_KERNEL_FBM, &kind) || kind == DGP_KERNEL_COMPOSITE ||
        !as_int(p[pos + 1], 0, 1, &gm) ||
        !as_int(p[pos + 2], 0, DGP_MAX_HYP, &nf) || pos + 3 + nf > n)
      return ctx->fail(DGP_ERR_ARG, "composite kernel: malformed base kernel %d", f);
    const int df = comp->factor_d[f], d0 = dev.dim0[f];
    const bool sliced = df != d;
    dgp_kernel_spec fs;
    fs.kind = kind;  fs.n_hyp = nf;  fs.grad_mode = gm;  fs.reserved = 0;  fs.hyp = p + pos + 3;  fs.noise = 0.0;
    DGP_TRY(parse_simple(ctx, &fs, df, &comp->factor[f]));
    pos += 3 + nf;
    total_hyp += nf;
    const Spec &sf = comp->factor[f];
    dev.cp[f] = make_cov_params(sf, false, df);
    if (kind_is_inner_product(kind)) {
      if (sliced)
        return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "composite kernel: Polynomial / FBM on a slice of the input");
      dev.any_ip = 1;
    } else if (kind == DGP_KERNEL_ARD_SE || sliced) {
      if (dp > MAX_ARD_DIM)
        return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "composite kernel: ARD-SE or sliced base kernels need d <= %d",
                         MAX_ARD_DIM);
      double w[MAX_ARD_DIM];
      for (int c = 0; c < MAX_ARD_DIM; ++c) w[c] = 0.0;
      for (int c = 0; c < df; ++c)
        w[d0 + c] = kind == DGP_KERNEL_ARD_SE ? 1.0 / (sf.hyp[1 + c] * sf.hyp[1 + c]) : 1.0;
      const int s = take_slot(w, kind_is_l1(kind) ? 1 : 0);
      if (s < 0)
        return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "composite kernel: more than %d distinct ARD-SE / sliced metrics",
                         MAX_ARD_SLOTS);
      dev.metric[f] = 2 + s;
      dev.ard_exact[f] = kind == DGP_KERNEL_ARD_SE && gm == DGP_GRAD_EXACT;
    } else if (kind_is_l1(kind)) {
      dev.metric[f] = 1;
      dev.any_l1 = 1;
    }
    dev.moment_off[f] = moments;
    moments += grad_partials_per_tile(dev.cp[f], dp);
  }
  dev.nslots = slots;
  dev.moment_off[F] = moments;
  for (int f = F + 1; f <= MAXF; ++f) dev.moment_off[f] = moments;
  if (total_hyp > DGP_MAX_HYP) return ctx->fail(DGP_ERR_ARG, "composite kernel: too many hyper-parameters");
  for (int t = 0; t < MAXT; ++t) {
    dev.coef[t] = 0.0;
    for (int f = 0; f < MAXF; ++f) dev.expo[t][f] = 0;
  }
  dev.simple = 1;
  for (int t = 0; t < MAXT; ++t) dev.mask[t] = 0u;
  for (int t = 0; t < T; ++t) {
    if (pos + 1 + F > n) return ctx->fail(DGP_ERR_ARG, "composite kernel: malformed term %d", t);
    dev.coef[t] = p[pos];
    for (int f = 0; f < F; ++f) {
      int e = 0;
      if (!as_int(p[pos + 1 + f], 0, 8, &e)) return ctx->fail(DGP_ERR_ARG, "composite kernel: bad exponent");
      dev.expo[t][f] = (unsigned char)e;
      if (e > 1) dev.simple = 0;
      if (e) dev.mask[t] |= 1u << f;
    }
    pos += 1 + F;
  }
  if (pos != n && pos + 2 * F != n)
    return ctx->fail(DGP_ERR_ARG, "composite kernel: %d trailing values in the program", n - pos);
  out->kind = DGP_KERNEL_COMPOSITE;
  out->n_hyp = total_hyp;
  out->grad_mode = in->grad_mode;
  out->noise = in->noise;
  out->comp = comp;
  return DGP_OK;
}

}  // namespace

bool spec_uses_raw_points(const Spec &sp) { return sp.comp && sp.comp->dev.any_ip; }

int32_t parse_spec(Ctx *ctx, const dgp_kernel_spec *in, int d, Spec *out) {
  if (!in || !in->hyp) return ctx->fail(DGP_ERR_ARG, "kernel spec / hyp is NULL");
  if (in->kind == DGP_KERNEL_COMPOSITE) return parse_composite(ctx, in, d, out);
  if (kind_is_inner_product(in->kind)) {
    // the non-stationary kernels run on the composite evaluation path: wrap as a one-factor expression
    if (in->n_hyp < 0 || in->n_hyp > 2) return ctx->fail(DGP_ERR_ARG, "kernel kind %d: bad n_hyp", in->kind);
    double prog[2 + 3 + 2 + 2] = {1.0, 1.0, (double)in->kind, (double)in->grad_mode, (double)in->n_hyp};
    int pos = 5;
    for (int i = 0; i < in->n_hyp; ++i) prog[pos++] = in->hyp[i];
    prog[pos++] = 1.0;  // coefficient
    prog[pos++] = 1.0;  // exponent
    dgp_kernel_spec w = *in;
    w.kind = DGP_KERNEL_COMPOSITE;
    w.n_hyp = pos;
    w.hyp = prog;
    return parse_composite(ctx, &w, d, out);
  }
  return parse_simple(ctx, in, d, out);
}

namespace {

int32_t parse_simple(Ctx *ctx, const dgp_kernel_spec *in, int d, Spec *out) {
  int want = -1;
  switch (in->kind) {
    case DGP_KERNEL_SE: want = 2; break;
    case DGP_KERNEL_RBF: want = 1; break;
    case DGP_KERNEL_ARD_SE: want = d + 1; break;
    case DGP_KERNEL_MATERN: want = 2; break;
    case DGP_KERNEL_PERIODIC: want = 3; break;
    case DGP_KERNEL_DIRAC: want = 1; break;
    case DGP_KERNEL_CAUCHY: want = 1; break;
    case DGP_KERNEL_LAPLACIAN: want = 1; break;
    case DGP_KERNEL_RATQUAD: want = 2; break;
    case DGP_KERNEL_TSTUDENT: want = 1; break;
    case DGP_KERNEL_POLYNOMIAL: want = 2; break;
    case DGP_KERNEL_FBM: want = 1; break;
    default: return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "unsupported kernel kind %d", in->kind);
  }
  if (in->n_hyp != want || want > DGP_MAX_HYP)
    return ctx->fail(DGP_ERR_ARG, "kernel kind %d with d=%d needs %d hyper-parameters, got %d", in->kind, d, want,
                     in->n_hyp);
  out->kind = in->kind;
  out->n_hyp = in->n_hyp;
  out->grad_mode = in->grad_mode;
  out->noise = in->noise;
  for (int i = 0; i < in->n_hyp; ++i) out->hyp[i] = in->hyp[i];
  if (in->kind == DGP_KERNEL_MATERN) {
    int p = (int)floor(fabs(in->hyp[0]));
    if (p > MAX_MATERN_P) return ctx->fail(DGP_ERR_UNSUPPORTED_KERNEL, "Matern order p=%d > %d", p, MAX_MATERN_P);
  }
  return DGP_OK;
}

}  // namespace

namespace {

// Repack host points (either layout) into row-major n_p x dp with zero padding, subtracting
// `center` (d values, nullable).  Returns the largest squared norm of the packed points.
double pack_points(const double *X, int64_t n, int d, int layout, int64_t np, int dp, std::vector<double> &out,
                   const double *center = nullptr) {
  out.assign((size_t)np * dp, 0.0);
  if (layout == DGP_ROW_MAJOR) {
    for (int64_t i = 0; i < n; ++i)
      for (int c = 0; c < d; ++c) out[(size_t)i * dp + c] = X[(size_t)i * d + c] - (center ? center[c] : 0.0);
  } else {
    for (int c = 0; c < d; ++c)
      for (int64_t i = 0; i < n; ++i) out[(size_t)i * dp + c] = X[(size_t)c * n + i] - (center ? center[c] : 0.0);
  }
  double mx = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int c = 0; c < d; ++c) s += out[(size_t)i * dp + c] * out[(size_t)i * dp + c];
    if (s > mx || s != s) mx = s;
  }
  return mx;
}

// True when two of the first n packed rows are equal as points (x - y == 0 in every coordinate).
// Open-addressing table over a 64-bit row hash (-0.0 folded onto +0.0); equal hashes are compared exactly.
bool has_duplicate_rows(const std::vector<double> &pts, int64_t n, int dp) {
  size_t cap = 1;
  while (cap < (size_t)n * 2) cap <<= 1;
  std::vector<int64_t> slot(cap, -1);
  std::vector<uint64_t> hs(cap);
  for (int64_t i = 0; i < n; ++i) {
    uint64_t acc = 0x9e3779b97f4a7c15ull;
    for (int c = 0; c < dp; ++c) {
      const double v = pts[(size_t)i * dp + c] + 0.0;
      uint64_t b;
      memcpy(&b, &v, sizeof b);
      acc = (acc ^ b) * 0xff51afd7ed558ccdull;
      acc ^= acc >> 29;
    }
    size_t p = acc & (cap - 1);
    while (slot[p] >= 0) {
      if (hs[p] == acc) {
        const double *a = &pts[(size_t)i * dp], *q = &pts[(size_t)slot[p] * dp];
        bool same = true;
        for (int c = 0; c < dp && same; ++c) same = (a[c] - q[c] == 0.0);
        if (same) return true;
      }
      p = (p + 1) & (cap - 1);
    }
    slot[p] = i;
    hs[p] = acc;
  }
  return false;
}

// Lean assembly precondition for a training matrix with a noise term: distinct points stay distinct
// under the kernel's own input scaling (ARD-SE divides by b_c; extreme b_c could collapse them).
int points_stay_distinct(bool has_dup, const Spec &sp, int d) {
  if (has_dup) return 0;
  if (sp.kind == DGP_KERNEL_ARD_SE)
    for (int c = 0; c < d; ++c) {
      const double ib = fabs(1.0 / sp.hyp[1 + c]);
      if (!(ib >= 1e-100 && ib <= 1e100)) return 0;
    }
  return 1;
}

void column_means(const double *X, int64_t n, int d, int layout, std::vector<double> &mean) {
  mean.assign(d, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < d; ++c) mean[c] += (layout == DGP_ROW_MAJOR) ? X[(size_t)i * d + c] : X[(size_t)c * n + i];
  for (int c = 0; c < d; ++c) {
    mean[c] /= (double)n;
    if (!std::isfinite(mean[c])) mean[c] = 0.0;
  }
}

// Whether the Gram/DMMA assembly may be used for points whose largest squared norm is max_sq_norm
// (for ARD-SE the kernel sees the inputs scaled by 1/b_c: bound the scaled norm).
int pick_gram(const Ctx *ctx, const Spec &sp, const CovParams &cp, int dp, int d, double max_sq_norm) {
  if (ctx->assembly == 1) return 0;
  if (cp.kind != DGP_KERNEL_SE && cp.kind != DGP_KERNEL_RBF && cp.kind != DGP_KERNEL_ARD_SE &&
      cp.kind != DGP_KERNEL_MATERN)
    return 0;
  if (dp > 32) return 0;
  if (ctx->assembly >= 2) return 1;
  double scale = 1.0;
  if (sp.kind == DGP_KERNEL_ARD_SE) {
    scale = 0.0;
    for (int c = 0; c < d; ++c) {
      const double ib2 = 1.0 / (sp.hyp[1 + c] * sp.hyp[1 + c]);
      if (ib2 > scale || ib2 != ib2) scale = ib2;
    }
  }
  return gram_is_accurate(cp, dp, max_sq_norm * scale) ? 1 : 0;
}

struct ScaleVec {
  double v[128];
};

__global__ void scale_points_byval_kernel(const double *X, double *Xs, ScaleVec s, int64_t total, int dp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) Xs[i] = X[i] * s.v[i % dp];
}

// ARD-SE: Xs = X * diag(1/b).  No-op (returns X) for the other kinds.
int32_t scaled_points(Ctx *ctx, cudaStream_t st, const Spec &sp, const double *X, int64_t np, int dp, int d,
                      const char *bufname, const double **out) {
  if (sp.kind != DGP_KERNEL_ARD_SE) {
    *out = X;
    return DGP_OK;
  }
  if (dp > 128) return ctx->fail(DGP_ERR_SHAPE, "ARD-SE supports at most 128 features");
  DGP_BUF(Xs, double, ctx, bufname, (size_t)np * dp * sizeof(double));
  ScaleVec s;
  for (int c = 0; c < 128; ++c) s.v[c] = (c < d) ? 1.0 / sp.hyp[1 + c] : 0.0;
  int64_t total = np * dp;
  scale_points_byval_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, Xs, s, total, dp);
  ctx->launches++;
  DGP_CUDA_OK(ctx, cudaGetLastError());
  *out = Xs;
  return DGP_OK;
}

}  // namespace

int points_stay_distinct_pub(bool has_dup, const Spec &sp, int d) { return points_stay_distinct(has_dup, sp, d); }

int pick_gram_pub(const Ctx *ctx, const Spec &sp, const CovParams &cp, int dp, int d, double max_sq_norm) {
  return pick_gram(ctx, sp, cp, dp, d, max_sq_norm);
}

int32_t scaled_points_pub(Ctx *ctx, cudaStream_t st, const Spec &sp, const double *X, int64_t np, int dp, int d,
                          const char *bufname, const double **out) {
  return scaled_points(ctx, st, sp, X, np, dp, d, bufname, out);
}

namespace {

__global__ void pad_identity_kernel(double *A, int64_t ld, int64_t n, int64_t np) {
  // rows/cols >= n of the np x np matrix: identity
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= np * np) return;
  int64_t r = idx % np, c = idx / np;
  if (r >= n || c >= n) A[r + c * ld] = (r == c) ? 1.0 : 0.0;
}

void tick(Ctx *ctx, int slot, int i, cudaStream_t on = nullptr) {
  if (ctx->timing && slot == 0) cudaEventRecord(ctx->tev[i], on ? on : ctx->stream);
}

// misc_event slots: 2 * slot and 2 * slot + 1 (slot = 0..3) fork / join the solves of an evaluation slot
constexpr size_t EV_COV_READY = 14;   // dgp_predict: Sigma* complete on the main stream
constexpr size_t EV_OUT_READY = 15;   // copy_out_2d: everything enqueued so far

cudaEvent_t misc_event(Ctx *ctx, size_t i) {
  while (ctx->ev_misc.size() <= i) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    ctx->ev_misc.push_back(e);
  }
  return ctx->ev_misc[i];
}

// Device matrix (column-major, ld_src) -> host matrix in pageable memory (ld_dst).  A transfer into pageable memory
// is staged by the driver at a few GB/s; from OUT_MIN bytes up it goes instead through two pinned 16 MB buffers on a
// stream of its own, the DMA of one chunk overlapping the host copy of the previous one.  Waits for `ready` (an event
// on the producing stream; nullptr: everything enqueued on S so far) and returns when the data is in `dst`; work
// enqueued on S after `ready` keeps running underneath.
constexpr size_t OUT_CHUNK = (size_t)16 << 20, OUT_MIN = (size_t)32 << 20;

int32_t copy_out_2d(Ctx *ctx, cudaStream_t S, cudaEvent_t ready, double *dst, int64_t ld_dst, const double *src,
                    int64_t ld_src, int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return DGP_OK;
  if ((size_t)rows * cols * sizeof(double) < OUT_MIN || (size_t)rows * sizeof(double) > OUT_CHUNK) {
    if (ready) DGP_CUDA_OK(ctx, cudaStreamWaitEvent(S, ready, 0));
    DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(dst, ld_dst * sizeof(double), src, ld_src * sizeof(double), rows * sizeof(double),
                                       cols, cudaMemcpyDeviceToHost, S));
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
    return DGP_OK;
  }
  if (!ctx->out_stream) {
    DGP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      DGP_CUDA_OK(ctx, cudaMallocHost(&ctx->h_out[b], OUT_CHUNK));
      DGP_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->out_ev[b], cudaEventDisableTiming));
    }
  }
  cudaStream_t O = ctx->out_stream;
  if (!ready) {
    ready = misc_event(ctx, EV_OUT_READY);
    DGP_CUDA_OK(ctx, cudaEventRecord(ready, S));
  }
  DGP_CUDA_OK(ctx, cudaStreamWaitEvent(O, ready, 0));
  const int64_t cc = (int64_t)(OUT_CHUNK / ((size_t)rows * sizeof(double)));
  auto drain = [&](int64_t c0, int b) -> int32_t {
    DGP_CUDA_OK(ctx, cudaEventSynchronize(ctx->out_ev[b]));
    const int64_t n = (cols - c0 < cc) ? cols - c0 : cc;
    for (int64_t c = 0; c < n; ++c)
      memcpy(dst + (size_t)(c0 + c) * ld_dst, ctx->h_out[b] + (size_t)c * rows, (size_t)rows * sizeof(double));
    return DGP_OK;
  };
  int64_t prev = -1;
  int i = 0;
  for (int64_t c0 = 0; c0 < cols; c0 += cc, ++i) {
    const int b = i & 1;
    const int64_t n = (cols - c0 < cc) ? cols - c0 : cc;
    DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(ctx->h_out[b], rows * sizeof(double), src + (size_t)c0 * ld_src,
                                       ld_src * sizeof(double), rows * sizeof(double), n, cudaMemcpyDeviceToHost, O));
    DGP_CUDA_OK(ctx, cudaEventRecord(ctx->out_ev[b], O));
    if (prev >= 0) DGP_TRY(drain(prev, b ^ 1));
    prev = c0;
  }
  return drain(prev, (i - 1) & 1);
}

// Stream pair of an evaluation slot; slot 0 is the context's own pair.
int32_t slot_streams(Ctx *ctx, int slot, cudaStream_t *S, cudaStream_t *P) {
  if (slot == 0) {
    *S = ctx->stream;
    *P = ctx->panel;
    return DGP_OK;
  }
  if (!ctx->slot_S[slot]) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    DGP_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->slot_S[slot], cudaStreamNonBlocking, lo));
    DGP_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->slot_P[slot], cudaStreamNonBlocking, hi));
    DGP_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->slot_ev[slot], cudaEventDisableTiming));
  }
  *S = ctx->slot_S[slot];
  *P = ctx->slot_P[slot];
  return DGP_OK;
}

// C += Psi1 Psi2'  (rank-basis_mp DMMA update; GPBasisFuncRegressionModel.scala:91-109)
int32_t add_basis_term(Ctx *ctx, cudaStream_t S, const double *psi1, int64_t ld1, const double *psi2, int64_t ld2,
                       int mp, double *Cm, int64_t ldc, int64_t rows_p, int64_t cols_p, bool lower) {
  GemmArgs g;
  g.A = psi1;  g.lda = ld1;  g.a_kmajor = 0;
  g.B = psi2;  g.ldb = ld2;  g.b_kmajor = 0;
  g.C = Cm;  g.ldc = ldc;
  g.M = (int)rows_p;  g.N = (int)cols_p;  g.K = mp;
  g.alpha = 1.0;  g.beta = 1.0;  g.lower_only = lower ? 1 : 0;
  return gemm(ctx, S, g);
}

// Upload row-major n x m features as column-major np x mp, zero padded.
int32_t upload_basis(Ctx *ctx, const double *psi, int64_t n, int64_t np, int m, int mp, double **out) {
  std::vector<double> h((size_t)np * mp, 0.0);
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < m; ++c) h[(size_t)c * np + i] = psi[(size_t)i * m + c];
  double *d = nullptr;
  if (cudaMalloc(&d, h.size() * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(DGP_ERR_OOM, "basis features: device allocation failed");
  }
  cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *out = d;
  return DGP_OK;
}

// Enqueue one energy (and optionally gradient-moment) evaluation on the streams and workspace of
// `slot`.  Results land in d_scal[0] = r' K^-1 r, d_scal[1] = sum log L_ii, d_scal[2..] = gradient
// moments; *d_info != 0 on a failed pivot.  Nothing here synchronises with the host.
int32_t enqueue_eval(Ctx *ctx, int slot, const Data *D, const Spec &sp, bool want_grad, double *d_scal, int *d_info,
                     int *ns_out) {
  cudaStream_t S, P;
  DGP_TRY(slot_streams(ctx, slot, &S, &P));
  const int64_t np = D->Np;
  char nm[6][32];
  const char *base[6] = {"K", "inv", "x", "Xs", "potrf_panel", "W"};
  for (int i = 0; i < 6; ++i) {
    if (slot == 0) snprintf(nm[i], sizeof nm[i], "%s", base[i]);
    else snprintf(nm[i], sizeof nm[i], "%s#%d", base[i], slot);
  }
  DGP_BUF(K, double, ctx, nm[0], (size_t)np * np * sizeof(double));
  DGP_BUF(inv, double, ctx, nm[1], (size_t)np * TILE * sizeof(double));
  DGP_BUF(x, double, ctx, nm[2], (size_t)np * sizeof(double));
  const double *Xk = nullptr;
  const double *Xbase = spec_uses_raw_points(sp) ? D->Xraw : D->X;
  DGP_TRY(scaled_points(ctx, S, sp, Xbase, np, D->dp, D->d, nm[3], &Xk));
  CovParams cp = make_cov_params(sp, true, D->d);

  tick(ctx, slot, 0);
  {
    NvtxRange r("dgp:assemble");
    AssembleArgs a;
    a.X1 = Xk;  a.X2 = Xk;  a.dp = D->dp;
    a.rows = D->N;  a.cols = D->N;  a.rows_p = np;  a.cols_p = np;
    a.out = K;  a.ld = np;
    a.symmetric = 1;  a.lower_only = 1;
    a.use_gram = pick_gram(ctx, sp, cp, D->dp, D->d, D->max_sq_norm);
    a.nodup = points_stay_distinct(D->has_dup, sp, D->d);
    DGP_TRY(assemble(ctx, S, cp, a));
    if (D->psi) DGP_TRY(add_basis_term(ctx, S, D->psi, np, D->psi, np, D->basis_mp, K, np, np, np, true));
  }
  tick(ctx, slot, 1);
  {
    NvtxRange r("dgp:potrf");
    // The factorisation may go through a CUDA graph (option "graphs"; one per slot and size): its status word must
    // then sit at a fixed address, so it is per slot and copied to the candidate's word afterwards.
    DGP_BUF(slot_info, int, ctx, "slot_info", 16 * sizeof(int));
    int *sinfo = slot_info + slot;
    DGP_CUDA_OK(ctx, cudaMemsetAsync(sinfo, 0, sizeof(int), S));
    DGP_TRY(potrf_blocked_graph(ctx, S, ctx->lookahead ? P : S, K, np, np, inv, sinfo, nm[4]));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(d_info, sinfo, sizeof(int), cudaMemcpyDeviceToDevice, S));
  }
  tick(ctx, slot, 2);
  // The two triangular solves are a chain of 2 N/128 small dependent launches that leaves the GPU almost idle
  // (5 ms at N = 32768); with a gradient to compute they run on the panel stream underneath the triangular
  // inverse, which needs only L.
  cudaStream_t Z = (want_grad && P != S) ? P : S;
  if (Z != S) {
    cudaEvent_t e = misc_event(ctx, 2 * slot);
    DGP_CUDA_OK(ctx, cudaEventRecord(e, S));
    DGP_CUDA_OK(ctx, cudaStreamWaitEvent(Z, e, 0));
  }
  {
    NvtxRange r("dgp:solve");
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(x, D->r, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice, Z));
    DGP_TRY(trsv_lower_fwd(ctx, Z, K, np, inv, np, x));
    DGP_TRY(trsv_lower_bwd(ctx, Z, K, np, inv, np, x));
    DGP_TRY(dot(ctx, Z, D->r, x, np, d_scal + 0));
    DGP_TRY(diag_log_sum(ctx, Z, K, np, np, d_scal + 1));
  }
  tick(ctx, slot, 3, Z);
  if (want_grad) {
    DGP_BUF(W, double, ctx, nm[5], (size_t)np * np * sizeof(double));
    DGP_BUF(Kinv, double, ctx, "Kinv", (size_t)np * np * sizeof(double));
    {
      NvtxRange r("dgp:inverse");
      if (ctx->inverse_transposed) {  // W holds U = L^-T
        DGP_TRY(trtri_lower_transposed_on(ctx, S, K, np, inv, np, W, np, Kinv, np));
        DGP_TRY(lauum_upper_on(ctx, S, W, np, np, Kinv, np));
      } else {
        DGP_TRY(trtri_lower_on(ctx, S, K, np, inv, np, W, np, Kinv, np));
        DGP_TRY(lauum_lower_on(ctx, S, W, np, np, Kinv, np));
      }
    }
    tick(ctx, slot, 4);
    if (Z != S) {  // the trace pass needs alpha
      cudaEvent_t e = misc_event(ctx, 2 * slot + 1);
      DGP_CUDA_OK(ctx, cudaEventRecord(e, Z));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(S, e, 0));
    }
    const int64_t mt = np / TILE, tiles = mt * (mt + 1) / 2;
    const int nsmax = grad_partials_per_tile(cp, D->dp);
    DGP_BUF(partials, double, ctx, "partials", (size_t)tiles * nsmax * sizeof(double));
    GradArgs g;
    g.X = Xbase;  g.Xs = Xk;  g.alpha = x;  g.Kinv = Kinv;  g.ld = np;
    g.n = D->N;  g.np = np;  g.dp = D->dp;
    g.partials = partials;  g.sums = d_scal + 2;
    // same conditions as the lean assembly of the strictly-lower tiles: Gram accuracy gate, pairwise distinct points
    g.lean = ctx->grad_lean && ctx->assembly != 3 && pick_gram(ctx, sp, cp, D->dp, D->d, D->max_sq_norm) &&
             points_stay_distinct(D->has_dup, sp, D->d);
    {
      NvtxRange r("dgp:grad_trace");
      DGP_TRY(grad_trace(ctx, S, cp, g, ns_out));
    }
    tick(ctx, slot, 5);
  }
  return DGP_OK;
}

double energy_from(const Data *D, double quad, double logsum) {
  // AbstractGPRegressionModel.scala:527-529: 0.5*(d + btrace(blog(Lmat)) + rows*log(2*Pi))
  return 0.5 * (quad + logsum + (double)D->N * log(2.0 * M_PI));
}

void collect_timings(Ctx *ctx, bool grad) {
  for (int i = 0; i < 16; ++i) ctx->timings[i] = 0.0;
  if (!ctx->timing) return;
  float ms;
  auto el = [&](int a, int b) { cudaEventElapsedTime(&ms, ctx->tev[a], ctx->tev[b]); return (double)ms; };
  ctx->timings[0] = el(0, 1);
  ctx->timings[1] = el(1, 2);
  ctx->timings[2] = el(2, 3);
  if (grad) {
    ctx->timings[3] = el(2, 4);  // the inverse starts with the solves, right after the Cholesky
    ctx->timings[4] = el(4, 5);
    ctx->timings[8] = el(0, 5);
  } else {
    ctx->timings[8] = el(0, 3);
  }
}

}  // namespace

struct Model {
  Spec spec;
  int64_t N = 0, Np = 0;
  int d = 0, dp = 0;
  double *X = nullptr;    // own copy of the (unscaled) padded training points
  double *Xk = nullptr;   // points the kernel consumes (ARD: scaled copy, else == X)
  double *L = nullptr;    // Np x Np
  double *inv = nullptr;  // Np x 128
  double *alpha = nullptr;
  double energy = 0.0, quad = 0.0, logdet_half = 0.0;
  std::vector<double> center;  // the training set's column means (test points are shifted alike)
  double max_sq_norm = 0.0;
  // optional covariance-only spec for the cross / test matrices (general noise models)
  bool has_cross = false;
  Spec cross_spec;
  double *Xk_cross = nullptr;  // training points as cross_spec consumes them (ARD: scaled copy, else == X)
  // GPBasisFuncRegressionModel: own copy of the training features, and the test features of the next predict
  double *psi = nullptr, *psi_s = nullptr;
  int basis_m = 0, basis_mp = 0;
  int64_t psi_s_n = 0;
};

}  // namespace dgp

struct dgp_model : dgp::Model {};

using namespace dgp;

extern "C" {

int32_t dgp_spec_grad_len(const dgp_kernel_spec *spec) {
  if (!spec || !spec->hyp || spec->n_hyp < 0) return -DGP_ERR_ARG;
  if (spec->kind != DGP_KERNEL_COMPOSITE) return spec->n_hyp + 1;
  const double *p = spec->hyp;
  const int n = spec->n_hyp;
  if (n < 2 || !(p[0] >= 1 && p[0] <= MAXF)) return -DGP_ERR_ARG;
  int pos = 2, total = 0;
  for (int f = 0; f < (int)p[0]; ++f) {
    if (pos + 3 > n || !(p[pos + 2] >= 0 && p[pos + 2] <= DGP_MAX_HYP)) return -DGP_ERR_ARG;
    total += (int)p[pos + 2];
    pos += 3 + (int)p[pos + 2];
  }
  return total + 1;
}

const char *dgp_version(void) { return "dgp-b200 0.1 (sm_100a, FP64 DMMA)"; }

int32_t dgp_ctx_create(int32_t device, dgp_ctx **out) {
  if (!out) return DGP_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DGP_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= count) return DGP_ERR_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DGP_ERR_CUDA;
  if (prop.major < 10) return DGP_ERR_CUDA;  // built for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return DGP_ERR_CUDA;
  dgp_ctx *ctx = new dgp_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  ctx->prio_lo = lo;
  ctx->prio_hi = hi;
  bool ok = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, lo) == cudaSuccess &&
            cudaStreamCreateWithPriority(&ctx->panel, cudaStreamNonBlocking, hi) == cudaSuccess &&
            cudaEventCreate(&ctx->ev_a) == cudaSuccess && cudaEventCreate(&ctx->ev_b) == cudaSuccess;
  for (int i = 0; i < 16 && ok; ++i) ok = cudaEventCreate(&ctx->tev[i]) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_info, 4096 * sizeof(int)) == cudaSuccess &&
       cudaMallocHost(&ctx->h_info, 4096 * sizeof(int)) == cudaSuccess &&
       cudaMallocHost(&ctx->h_scal, 4096 * 160 * sizeof(double)) == cudaSuccess;
  if (ok) {
    // dgp_data blocks come from a pool of the library's own, which keeps freed blocks instead of returning them to
    // the driver at every synchronisation (a create / destroy pair per evaluation is the e2e pattern); the device's
    // default pool, shared with every other cudaMallocAsync user of the process, is left alone.
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    uint64_t keep = UINT64_MAX;
    ok = cudaMemPoolCreate(&ctx->data_pool, &props) == cudaSuccess &&
         cudaMemPoolSetAttribute(ctx->data_pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
  }
  if (ok) ok = gemm_init(ctx) == DGP_OK && potrf_init(ctx) == DGP_OK && solve_init(ctx) == DGP_OK;
  if (!ok) {
    delete ctx;
    return DGP_ERR_CUDA;
  }
  for (int i = 0; i < 16; ++i) ctx->timings[i] = 0.0;
  *out = ctx;
  return DGP_OK;
}

int32_t dgp_ctx_destroy(dgp_ctx *ctx) {
  if (!ctx) return DGP_OK;
  cudaSetDevice(ctx->device);
  dist_destroy(ctx);
  ctx->drop_graphs();
  gemm_release(ctx);
  ctx->release_all();
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto e : ctx->ev_misc) cudaEventDestroy(e);
  for (int i = 0; i < 16; ++i) cudaEventDestroy(ctx->tev[i]);
  cudaEventDestroy(ctx->ev_a);
  cudaEventDestroy(ctx->ev_b);
  cudaFree(ctx->d_info);
  cudaFreeHost(ctx->h_info);
  cudaFreeHost(ctx->h_scal);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  for (int b = 0; b < 2; ++b) {
    if (ctx->h_out[b]) cudaFreeHost(ctx->h_out[b]);
    if (ctx->out_ev[b]) cudaEventDestroy(ctx->out_ev[b]);
  }
  if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
  if (ctx->data_pool) cudaMemPoolDestroy(ctx->data_pool);
  for (int i = 1; i < 4; ++i) {
    if (ctx->slot_S[i]) cudaStreamDestroy(ctx->slot_S[i]);
    if (ctx->slot_P[i]) cudaStreamDestroy(ctx->slot_P[i]);
    if (ctx->slot_ev[i]) cudaEventDestroy(ctx->slot_ev[i]);
  }
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->panel);
  delete ctx;
  return DGP_OK;
}

int32_t dgp_ctx_trim(dgp_ctx *ctx) {
  if (!ctx) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  ctx->release_all();  // synchronises the device first
  gemm_release(ctx);
  if (ctx->data_pool) cudaMemPoolTrimTo(ctx->data_pool, 0);
  return DGP_OK;
}

int32_t dgp_ctx_info(dgp_ctx *ctx, int64_t *out, int32_t n) {
  if (!ctx || !out || n < 4) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  size_t fr = 0, tot = 0;
  DGP_CUDA_OK(ctx, cudaMemGetInfo(&fr, &tot));
  out[0] = ctx->device;
  out[1] = ctx->sm_count;
  out[2] = (int64_t)tot;
  out[3] = (int64_t)fr;
  return DGP_OK;
}

const char *dgp_last_error(const dgp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t dgp_ctx_set_option(dgp_ctx *ctx, const char *key, int64_t value) {
  if (!ctx || !key) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  ctx->drop_graphs();  // every tunable below changes what a captured factorisation would launch
  if (!strcmp(key, "nb")) {
    if (value < 0 || value % TILE) return ctx->fail(DGP_ERR_ARG, "nb must be 0 (auto) or a positive multiple of 128");
    ctx->nb = value;
  } else if (!strcmp(key, "lookahead")) {
    ctx->lookahead = value != 0;
  } else if (!strcmp(key, "gemm_handover")) {
    if (value < 0 || value > 3) return ctx->fail(DGP_ERR_ARG, "gemm_handover must be in 0..3");
    ctx->gemm_handover = (int)value;
  } else if (!strcmp(key, "gemm_variant")) {
    if (value != 2 && value != 4) return ctx->fail(DGP_ERR_ARG, "gemm_variant must be 2 (cp.async) or 4 (bulk copy)");
    ctx->gemm_variant = (int)value;
  } else if (!strcmp(key, "gemm_v4_bk32_mask")) {
    if (value < 0 || value > 15) return ctx->fail(DGP_ERR_ARG, "gemm_v4_bk32_mask must be in 0..15");
    ctx->gemm_v4_bk32_mask = (int)value;
  } else if (!strcmp(key, "gemm_v4_mask")) {
    if (value < 0 || value > 15) return ctx->fail(DGP_ERR_ARG, "gemm_v4_mask must be in 0..15");
    ctx->gemm_v4_mask = (int)value;
  } else if (!strcmp(key, "graphs")) {
    ctx->use_graphs = value != 0;
  } else if (!strcmp(key, "leaf_variant")) {
    if (value < 0 || value > 1) return ctx->fail(DGP_ERR_ARG, "leaf_variant must be 0 or 1");
    ctx->leaf_variant = (int)value;
  } else if (!strcmp(key, "trsv_chained")) {
    ctx->trsv_chained = value != 0;
  } else if (!strcmp(key, "inverse_transposed")) {
    ctx->inverse_transposed = value != 0;
  } else if (!strcmp(key, "gemm_no_prefetch")) {
    ctx->gemm_no_prefetch = value != 0;
  } else if (!strcmp(key, "gemm_bk")) {
    if (value != 16 && value != 32) return ctx->fail(DGP_ERR_ARG, "gemm_bk must be 16 or 32");
    ctx->gemm_bk = (int)value;
  } else if (!strcmp(key, "assembly")) {
    if (value < 0 || value > 3)
      return ctx->fail(DGP_ERR_ARG, "assembly must be 0 (auto), 1 (exact differences), 2 (Gram) or 3 (checked Gram)");
    ctx->assembly = (int)value;
  } else if (!strcmp(key, "grad_lean")) {
    if (value < 0 || value > 1) return ctx->fail(DGP_ERR_ARG, "grad_lean must be 0 (exact differences) or 1 (lean Gram pass)");
    ctx->grad_lean = (int)value;
  } else if (!strcmp(key, "assembly_store")) {
    if (value < 0 || value > 1) return ctx->fail(DGP_ERR_ARG, "assembly_store must be 0 (direct) or 1 (bulk)");
    ctx->assembly_store = (int)value;
  } else if (!strcmp(key, "batch_streams")) {
    if (value < 0 || value > 4) return ctx->fail(DGP_ERR_ARG, "batch_streams must be in 0..4 (0 = automatic)");
    ctx->batch_streams = (int)value;
  } else if (!strcmp(key, "timing")) {
    ctx->timing = value != 0;
  } else {
    return ctx->fail(DGP_ERR_ARG, "unknown option '%s'", key);
  }
  return DGP_OK;
}

int32_t dgp_get_timings(const dgp_ctx *ctx, double *ms, int32_t n) {
  if (!ctx || !ms) return DGP_ERR_ARG;
  for (int i = 0; i < n; ++i) ms[i] = i < 16 ? ctx->timings[i] : 0.0;
  return DGP_OK;
}

// ---- data ------------------------------------------------------------------------------------
int32_t dgp_data_create(dgp_ctx *ctx, const double *X, int64_t N, int32_t d, int32_t x_layout, const double *y,
                        const double *mean, dgp_data **out) {
  if (!ctx || !X || !y || !out) return DGP_ERR_ARG;
  if (N <= 0 || d <= 0) return ctx->fail(DGP_ERR_SHAPE, "data: N=%lld d=%d", (long long)N, d);
  cudaSetDevice(ctx->device);
  dgp_data *D = new dgp_data();
  D->N = N;
  D->Np = round_up(N, TILE);
  D->d = d;
  D->dp = (int)round_up(d, 4);
  // one device block [ centred X | raw X | r ] from the stream-ordered pool, filled by one copy out of a
  // pinned staging buffer: a create / destroy pair costs host packing only (cudaMalloc / cudaFree took ~2 ms each)
  // (+ TILE zero rows behind each point set: dgp_data_kernel_block assembles blocks that start at any row)
  const size_t nx = (size_t)(D->Np + TILE) * D->dp, nr = (size_t)D->Np, total = 2 * nx + nr;
  if (ctx->h_stage_bytes < total * sizeof(double)) {
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr;
    ctx->h_stage_bytes = 0;
    if (cudaMallocHost(&ctx->h_stage, total * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      delete D;
      return ctx->fail(DGP_ERR_OOM, "data: pinned staging allocation failed");
    }
    ctx->h_stage_bytes = total * sizeof(double);
  }
  double *hs = (double *)ctx->h_stage;
  std::vector<double> hx;
  column_means(X, N, d, x_layout, D->center);
  D->max_sq_norm = pack_points(X, N, d, x_layout, D->Np, D->dp, hx, D->center.data());
  D->has_dup = has_duplicate_rows(hx, N, D->dp);
  memset(hs, 0, 2 * nx * sizeof(double));
  memcpy(hs, hx.data(), hx.size() * sizeof(double));
  pack_points(X, N, d, x_layout, D->Np, D->dp, hx, nullptr);
  memcpy(hs + nx, hx.data(), hx.size() * sizeof(double));
  for (int64_t i = 0; i < D->Np; ++i) hs[2 * nx + i] = i < N ? y[i] - (mean ? mean[i] : 0.0) : 0.0;
  if (cudaMallocFromPoolAsync((void **)&D->block, total * sizeof(double), ctx->data_pool, ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    delete D;
    return ctx->fail(DGP_ERR_OOM, "data: device allocation failed");
  }
  D->X = D->block;
  D->Xraw = D->block + nx;
  D->r = D->block + 2 * nx;
  cudaMemcpyAsync(D->block, hs, total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *out = D;
  return DGP_OK;
}

int32_t dgp_data_destroy(dgp_ctx *ctx, dgp_data *D) {
  if (!D) return DGP_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (ctx) cudaFreeAsync(D->block, ctx->stream);
  else cudaFree(D->block);
  cudaFree(D->psi);
  delete D;
  return DGP_OK;
}

int32_t dgp_data_set_basis(dgp_ctx *ctx, dgp_data *D, const double *psi, int32_t m) {
  if (!ctx || !D || m < 0 || (m > 0 && !psi)) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(D->psi);
  D->psi = nullptr;
  D->basis_m = D->basis_mp = 0;
  if (m == 0) return DGP_OK;
  const int mp = (int)round_up(m, 32);
  DGP_TRY(upload_basis(ctx, psi, D->N, D->Np, m, mp, &D->psi));
  D->basis_m = m;
  D->basis_mp = mp;
  return DGP_OK;
}

// ---- energy / gradient -------------------------------------------------------------------------
static int32_t energy_impl(dgp_ctx *ctx, const dgp_data *D, const dgp_kernel_spec *spec, double *e, double *g) {
  if (!ctx || !D || !spec) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, D->d, &sp));
  DGP_BUF(d_scal, double, ctx, "scal", 4096 * 160 * sizeof(double));
  const bool want_grad = g != nullptr;
  int ns = 0;
  int32_t st = enqueue_eval(ctx, 0, D, sp, want_grad, d_scal, ctx->d_info, &ns);
  if (st != DGP_OK) {
    cudaStreamSynchronize(ctx->stream);
    return st;
  }
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scal, d_scal, 160 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  collect_timings(ctx, want_grad);
  if (ctx->h_info[0] != 0) {
    if (e) *e = INFINITY;
    if (g)
      for (int i = 0; i <= sp.n_hyp; ++i) g[i] = nan("");
    return ctx->fail(DGP_ERR_NOT_PD, "Cholesky failed: pivot %d is not positive", ctx->h_info[0] - 1);
  }
  if (e) *e = energy_from(D, ctx->h_scal[0], ctx->h_scal[1]);
  if (g) grad_finish(sp, D->d, ctx->h_scal + 2, ns, g);
  return DGP_OK;
}

int32_t dgp_energy(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, double *e) {
  if (!e) return DGP_ERR_ARG;
  return energy_impl(ctx, data, spec, e, nullptr);
}

int32_t dgp_energy_terms(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, double *quad,
                         double *logdet_half) {
  if (!quad || !logdet_half) return DGP_ERR_ARG;
  double e = 0.0;
  const int32_t st = energy_impl(ctx, data, spec, &e, nullptr);
  if (st != DGP_OK) {
    *quad = *logdet_half = nan("");
    return st;
  }
  *quad = ctx->h_scal[0];  // left in the pinned read-back buffer by energy_impl
  *logdet_half = ctx->h_scal[1];
  return DGP_OK;
}

int32_t dgp_energy_grad(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, double *e, double *g) {
  if (!g) return DGP_ERR_ARG;
  return energy_impl(ctx, data, spec, e, g);
}

int32_t dgp_energy_batch(dgp_ctx *ctx, const dgp_data *D, const dgp_kernel_spec *specs, int32_t n, double *e,
                         int32_t *status) {
  if (!ctx || !D || !specs || !e || !status || n < 0) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  DGP_BUF(d_scal, double, ctx, "scal", 4096 * 160 * sizeof(double));
  // Candidates are enqueued back to back in chunks of 4096, round-robin over `batch_streams`
  // evaluation slots (each with its own stream pair and workspace): the GPU never waits for the
  // host between candidates, and the panel chain of one candidate overlaps the trailing updates of
  // another.
  // automatic choice (0): four below N = 12288 (N = 8192: 164 against 158 candidates/s with three), three above (N = 16384:
  // no difference, and a slot holds an N x N workspace); scripts/bench_batch_nb.py
  int nslots = ctx->batch_streams == 0 ? (D->Np <= 12288 ? 4 : 3) : ctx->batch_streams;
  nslots = nslots < 1 ? 1 : (nslots > 4 ? 4 : nslots);
  if (n < nslots) nslots = n > 0 ? n : 1;
  double batch_ms = 0.0;
  for (int32_t base = 0; base < n; base += 4096) {
    DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));  // every slot's stream starts behind the main stream
    const int32_t cnt = (n - base < 4096) ? n - base : 4096;
    std::vector<int32_t> pre(cnt, DGP_OK);
    for (int32_t i = 0; i < cnt; ++i) {
      Spec sp;
      int ns = 0;
      int32_t st = parse_spec(ctx, &specs[base + i], D->d, &sp);
      if (st == DGP_OK && i < nslots && i > 0) {  // first use of a slot in this chunk: order it behind ev_a
        cudaStream_t S1, P1;
        st = slot_streams(ctx, i % nslots, &S1, &P1);
        if (st == DGP_OK) cudaStreamWaitEvent(S1, ctx->ev_a, 0);
      }
      if (st == DGP_OK) st = enqueue_eval(ctx, i % nslots, D, sp, false, d_scal + (size_t)i * 160, ctx->d_info + i, &ns);
      pre[i] = st;
      if (st == DGP_ERR_CUDA || st == DGP_ERR_OOM) {
        cudaDeviceSynchronize();
        return st;
      }
    }
    for (int sl = 1; sl < nslots; ++sl) {  // the read-back on the main stream follows every slot
      if (!ctx->slot_S[sl]) continue;
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->slot_ev[sl], ctx->slot_S[sl]));
      DGP_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->slot_ev[sl], 0));
    }
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scal, d_scal, (size_t)cnt * 160 * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
      batch_ms += ms;
    }
    for (int32_t i = 0; i < cnt; ++i) {
      if (pre[i] != DGP_OK) {
        status[base + i] = pre[i];
        e[base + i] = nan("");
      } else if (ctx->h_info[i] != 0) {
        status[base + i] = DGP_ERR_NOT_PD;
        e[base + i] = INFINITY;
      } else {
        status[base + i] = DGP_OK;
        e[base + i] = energy_from(D, ctx->h_scal[(size_t)i * 160], ctx->h_scal[(size_t)i * 160 + 1]);
      }
    }
  }
  collect_timings(ctx, false);
  ctx->timings[9] = batch_ms;  // device time of the whole batch (first enqueue to the last read-back)
  return DGP_OK;
}

int32_t dgp_energy_batch_multi(dgp_ctx *const *ctxs, const dgp_data *const *data, int32_t n_ctx,
                               const dgp_kernel_spec *specs, int32_t n, double *e, int32_t *status) {
  if (!ctxs || !data || n_ctx < 1 || !specs || !e || !status || n < 0) return DGP_ERR_ARG;
  for (int32_t c = 0; c < n_ctx; ++c)
    if (!ctxs[c] || !data[c]) return DGP_ERR_ARG;
  if (n_ctx == 1) return dgp_energy_batch(ctxs[0], data[0], specs, n, e, status);
  for (int32_t c = 0; c < n_ctx; ++c) {
    if (data[c]->N != data[0]->N || data[c]->d != data[0]->d)
      return ctxs[0]->fail(DGP_ERR_SHAPE, "energy_batch_multi: the data handles must hold the same training set");
    for (int32_t b = 0; b < c; ++b)
      if (ctxs[b] == ctxs[c]) return ctxs[0]->fail(DGP_ERR_ARG, "energy_batch_multi: a context is listed twice");
  }
  // Candidate i runs on context i mod n_ctx (the round-robin of SURVEY 8e).  One host thread per context: a dgp_ctx
  // is single-threaded but distinct contexts are independent, and every entry point selects its own device, so the
  // shards run concurrently on their GPUs; only scalars come back.
  std::vector<std::vector<dgp_kernel_spec>> shard(n_ctx);
  std::vector<std::vector<double>> se(n_ctx);
  std::vector<std::vector<int32_t>> ss(n_ctx);
  for (int32_t i = 0; i < n; ++i) shard[i % n_ctx].push_back(specs[i]);
  std::vector<int32_t> rc(n_ctx, DGP_OK);
  std::vector<std::thread> workers;
  for (int32_t c = 0; c < n_ctx; ++c) {
    se[c].resize(shard[c].size());
    ss[c].resize(shard[c].size());
    workers.emplace_back([&, c] {
      rc[c] = shard[c].empty() ? DGP_OK
                               : dgp_energy_batch(ctxs[c], data[c], shard[c].data(), (int32_t)shard[c].size(), se[c].data(),
                                                  ss[c].data());
    });
  }
  for (auto &w : workers) w.join();
  for (int32_t c = 0; c < n_ctx; ++c)
    if (rc[c] != DGP_OK) {
      if (c != 0) ctxs[0]->fail(rc[c], "energy_batch_multi: context %d: %s", c, ctxs[c]->err.c_str());
      return rc[c];
    }
  for (int32_t i = 0; i < n; ++i) {
    e[i] = se[i % n_ctx][i / n_ctx];
    status[i] = ss[i % n_ctx][i / n_ctx];
  }
  return DGP_OK;
}

// ---- fit / predict ------------------------------------------------------------------------------
int32_t dgp_model_destroy(dgp_ctx *ctx, dgp_model *m) {
  if (!m) return DGP_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  if (m->Xk != m->X) cudaFree(m->Xk);
  cudaFree(m->X);
  cudaFree(m->psi);
  cudaFree(m->psi_s);
  if (m->Xk_cross && m->Xk_cross != m->X) cudaFree(m->Xk_cross);
  cudaFree(m->L);
  cudaFree(m->inv);
  cudaFree(m->alpha);
  delete m;
  return DGP_OK;
}

int32_t dgp_fit(dgp_ctx *ctx, const dgp_data *D, const dgp_kernel_spec *spec, dgp_model **out) {
  if (!ctx || !D || !spec || !out) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  *out = nullptr;
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, D->d, &sp));
  cudaStream_t S = ctx->stream;
  const int64_t np = D->Np;
  dgp_model *m = new dgp_model();
  m->spec = sp;
  m->N = D->N;  m->Np = np;  m->d = D->d;  m->dp = D->dp;
  m->center = D->center;  m->max_sq_norm = D->max_sq_norm;
  const bool raw = spec_uses_raw_points(sp);
  if (raw) m->center.assign(D->center.size(), 0.0);  // test points are then taken as they are
  bool ok = cudaMalloc(&m->X, (size_t)np * D->dp * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&m->L, (size_t)np * np * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&m->inv, (size_t)np * TILE * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&m->alpha, (size_t)np * sizeof(double)) == cudaSuccess;
  if (ok && sp.kind == DGP_KERNEL_ARD_SE) ok = cudaMalloc(&m->Xk, (size_t)np * D->dp * sizeof(double)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    dgp_model_destroy(ctx, m);
    return ctx->fail(DGP_ERR_OOM, "fit: device allocation failed for N=%lld", (long long)D->N);
  }
  auto bail = [&](int32_t st) {
    cudaStreamSynchronize(S);
    dgp_model_destroy(ctx, m);
    return st;
  };
  cudaMemcpyAsync(m->X, raw ? D->Xraw : D->X, (size_t)np * D->dp * sizeof(double), cudaMemcpyDeviceToDevice, S);
  if (sp.kind == DGP_KERNEL_ARD_SE) {
    ScaleVec s;
    for (int c = 0; c < 128; ++c) s.v[c] = (c < D->d) ? 1.0 / sp.hyp[1 + c] : 0.0;
    int64_t total = np * D->dp;
    scale_points_byval_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S>>>(m->X, m->Xk, s, total, D->dp);
  } else {
    m->Xk = m->X;
  }
  CovParams cp = make_cov_params(sp, true, D->d);
  DGP_BUF(d_scal, double, ctx, "scal", 4096 * 160 * sizeof(double));
  tick(ctx, 0, 0);
  AssembleArgs a;
  a.X1 = m->Xk;  a.X2 = m->Xk;  a.dp = D->dp;
  a.rows = D->N;  a.cols = D->N;  a.rows_p = np;  a.cols_p = np;
  a.out = m->L;  a.ld = np;  a.symmetric = 1;  a.lower_only = 1;
  a.use_gram = pick_gram(ctx, sp, cp, D->dp, D->d, D->max_sq_norm);
  a.nodup = points_stay_distinct(D->has_dup, sp, D->d);
  int32_t st = assemble(ctx, S, cp, a);
  if (st != DGP_OK) return bail(st);
  if (D->psi) {
    ok = cudaMalloc(&m->psi, (size_t)np * D->basis_mp * sizeof(double)) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      return bail(ctx->fail(DGP_ERR_OOM, "fit: device allocation failed"));
    }
    cudaMemcpyAsync(m->psi, D->psi, (size_t)np * D->basis_mp * sizeof(double), cudaMemcpyDeviceToDevice, S);
    m->basis_m = D->basis_m;
    m->basis_mp = D->basis_mp;
    st = add_basis_term(ctx, S, m->psi, np, m->psi, np, m->basis_mp, m->L, np, np, np, true);
    if (st != DGP_OK) return bail(st);
  }
  tick(ctx, 0, 1);
  cudaMemsetAsync(ctx->d_info, 0, sizeof(int), S);
  st = potrf_blocked(ctx, m->L, np, np, m->inv, ctx->d_info);
  if (st != DGP_OK) return bail(st);
  tick(ctx, 0, 2);
  cudaMemcpyAsync(m->alpha, D->r, (size_t)np * sizeof(double), cudaMemcpyDeviceToDevice, S);
  st = trsv_lower_fwd(ctx, S, m->L, np, m->inv, np, m->alpha);
  if (st == DGP_OK) st = trsv_lower_bwd(ctx, S, m->L, np, m->inv, np, m->alpha);
  if (st == DGP_OK) st = dot(ctx, S, D->r, m->alpha, np, d_scal);
  if (st == DGP_OK) st = diag_log_sum(ctx, S, m->L, np, np, d_scal + 1);
  if (st != DGP_OK) return bail(st);
  tick(ctx, 0, 3);
  cudaMemcpyAsync(ctx->h_scal, d_scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, S);
  cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, S);
  cudaError_t ce = cudaStreamSynchronize(S);
  if (ce != cudaSuccess) return bail(ctx->fail(DGP_ERR_CUDA, "fit: %s", cudaGetErrorString(ce)));
  collect_timings(ctx, false);
  if (ctx->h_info[0] != 0)
    return bail(ctx->fail(DGP_ERR_NOT_PD, "Cholesky failed: pivot %d is not positive", ctx->h_info[0] - 1));
  m->energy = energy_from(D, ctx->h_scal[0], ctx->h_scal[1]);
  m->quad = ctx->h_scal[0];
  m->logdet_half = ctx->h_scal[1];
  *out = m;
  return DGP_OK;
}

int32_t dgp_model_energy(dgp_ctx *ctx, const dgp_model *m, double *e) {
  if (!ctx || !m || !e) return DGP_ERR_ARG;
  *e = m->energy;
  return DGP_OK;
}

int32_t dgp_model_terms(dgp_ctx *ctx, const dgp_model *m, double *quad, double *logdet_half) {
  if (!ctx || !m || !quad || !logdet_half) return DGP_ERR_ARG;
  *quad = m->quad;
  *logdet_half = m->logdet_half;
  return DGP_OK;
}

int32_t dgp_model_set_cross_spec(dgp_ctx *ctx, dgp_model *m, const dgp_kernel_spec *spec) {
  if (!ctx || !m || !spec) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, m->d, &sp));
  if (spec_uses_raw_points(sp) != spec_uses_raw_points(m->spec))
    return ctx->fail(DGP_ERR_ARG, "cross spec: inner-product kernels must appear in both specs or in neither");
  cudaStream_t S = ctx->stream;
  cudaStreamSynchronize(S);
  if (m->Xk_cross && m->Xk_cross != m->X) cudaFree(m->Xk_cross);
  m->Xk_cross = m->X;
  if (sp.kind == DGP_KERNEL_ARD_SE) {
    if (cudaMalloc(&m->Xk_cross, (size_t)m->Np * m->dp * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      m->Xk_cross = nullptr;
      return ctx->fail(DGP_ERR_OOM, "cross spec: device allocation failed");
    }
    ScaleVec sv;
    for (int c = 0; c < 128; ++c) sv.v[c] = (c < m->d) ? 1.0 / sp.hyp[1 + c] : 0.0;
    const int64_t total = m->Np * m->dp;
    scale_points_byval_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S>>>(m->X, m->Xk_cross, sv, total, m->dp);
    ctx->launches++;
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  }
  m->cross_spec = sp;
  m->has_cross = true;
  return DGP_OK;
}

int32_t dgp_model_solve(dgp_ctx *ctx, const dgp_model *m, const double *b, int32_t k, double *out) {
  if (!ctx || !m || !b || !out || k <= 0) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStream_t S = ctx->stream;
  const int64_t np = m->Np, n = m->N;
  DGP_BUF(x, double, ctx, "solve_x", (size_t)np * sizeof(double));
  for (int j = 0; j < k; ++j) {
    // Lmat.t \\ (Lmat \\ b_j) with the padded tail zero (the padding block of L is the identity)
    DGP_CUDA_OK(ctx, cudaMemsetAsync(x, 0, (size_t)np * sizeof(double), S));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(x, b + (size_t)j * n, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, S));
    DGP_TRY(trsv_lower_fwd(ctx, S, m->L, np, m->inv, np, x));
    DGP_TRY(trsv_lower_bwd(ctx, S, m->L, np, m->inv, np, x));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(out + (size_t)j * n, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, S));
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));  // b / out are pageable host memory reused per column
  }
  return DGP_OK;
}

int32_t dgp_model_cross_apply(dgp_ctx *ctx, const dgp_model *m, const double *Xs, int64_t M, int32_t x_layout,
                              const double *v, int32_t k, double *out) {
  if (!ctx || !m || !Xs || !v || !out || M <= 0 || k <= 0) return DGP_ERR_ARG;
  if (m->psi && (!m->psi_s || m->psi_s_n != M))
    return ctx->fail(DGP_ERR_ARG, "cross_apply: the model has basis functions; call dgp_model_set_test_basis first");
  cudaSetDevice(ctx->device);
  cudaStream_t S = ctx->stream;
  const int64_t np = m->Np, mp = round_up(M, TILE), n = m->N;
  const int dp = m->dp;
  std::vector<double> hx;
  double nmax = pack_points(Xs, M, m->d, x_layout, mp, dp, hx, m->center.data());
  if (m->max_sq_norm > nmax) nmax = m->max_sq_norm;
  DGP_BUF(Xt, double, ctx, "Xt", (size_t)mp * dp * sizeof(double));
  DGP_BUF(Ks, double, ctx, "Ks", (size_t)np * mp * sizeof(double));
  DGP_BUF(vin, double, ctx, "cross_v", (size_t)np * sizeof(double));
  DGP_BUF(vout, double, ctx, "cross_out", (size_t)mp * sizeof(double));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(Xt, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, S));
  const Spec &cs = m->has_cross ? m->cross_spec : m->spec;
  const double *Xtrain = m->has_cross ? m->Xk_cross : m->Xk;
  const double *Xtk = nullptr;
  DGP_TRY(scaled_points(ctx, S, cs, Xt, mp, dp, m->d, "Xts", &Xtk));
  CovParams cpx = make_cov_params(cs, false, m->d);  // K* carries no noise (:279-286)
  AssembleArgs a;
  a.X1 = Xtrain;  a.X2 = Xtk;  a.dp = dp;
  a.rows = n;  a.cols = M;  a.rows_p = np;  a.cols_p = mp;
  a.out = Ks;  a.ld = np;
  a.use_gram = pick_gram(ctx, cs, cpx, dp, m->d, nmax);
  DGP_TRY(assemble(ctx, S, cpx, a));
  if (m->psi) DGP_TRY(add_basis_term(ctx, S, m->psi, np, m->psi_s, mp, m->basis_mp, Ks, np, np, mp, false));
  for (int j = 0; j < k; ++j) {
    DGP_CUDA_OK(ctx, cudaMemsetAsync(vin, 0, (size_t)np * sizeof(double), S));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(vin, v + (size_t)j * n, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, S));
    DGP_TRY(gemv_t(ctx, S, Ks, np, np, mp, vin, nullptr, vout));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(out + (size_t)j * M, vout, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, S));
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  }
  return DGP_OK;
}

int32_t dgp_model_set_test_basis(dgp_ctx *ctx, dgp_model *m, const double *psi_s, int64_t n_test, int32_t mb) {
  if (!ctx || !m || !psi_s || n_test <= 0) return DGP_ERR_ARG;
  if (!m->psi || mb != m->basis_m)
    return ctx->fail(DGP_ERR_SHAPE, "test basis: the model was fitted with %d basis functions, got %d", m->basis_m, mb);
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(m->psi_s);
  m->psi_s = nullptr;
  m->psi_s_n = 0;
  DGP_TRY(upload_basis(ctx, psi_s, n_test, round_up(n_test, TILE), mb, m->basis_mp, &m->psi_s));
  m->psi_s_n = n_test;
  return DGP_OK;
}

int32_t dgp_predict(dgp_ctx *ctx, const dgp_model *m, const double *Xs, int64_t M, int32_t x_layout,
                    const double *mean_s, double *mean, double *cov, int64_t ldcov, double *lo, double *hi,
                    double sigma) {
  if (!ctx || !m || !Xs || !mean || M <= 0) return DGP_ERR_ARG;
  if (cov && ldcov < M) return ctx->fail(DGP_ERR_SHAPE, "predict: ldcov < M");
  if ((lo == nullptr) != (hi == nullptr)) return ctx->fail(DGP_ERR_ARG, "predict: lo and hi go together");
  cudaSetDevice(ctx->device);
  cudaStream_t S = ctx->stream;
  const int64_t np = m->Np, mp = round_up(M, TILE);
  const int dp = m->dp;
  const bool need_cov = cov != nullptr || lo != nullptr;

  if (m->psi && (!m->psi_s || m->psi_s_n != M))
    return ctx->fail(DGP_ERR_ARG, "predict: the model has basis functions; call dgp_model_set_test_basis for these "
                                  "%lld test points first", (long long)M);
  std::vector<double> hx;
  double nmax = pack_points(Xs, M, m->d, x_layout, mp, dp, hx, m->center.data());
  if (m->max_sq_norm > nmax) nmax = m->max_sq_norm;
  DGP_BUF(Xt, double, ctx, "Xt", (size_t)mp * dp * sizeof(double));
  DGP_BUF(Ks, double, ctx, "Ks", (size_t)np * mp * sizeof(double));
  DGP_BUF(dmean, double, ctx, "pmean", (size_t)mp * 3 * sizeof(double));
  double *dbase = dmean + mp, *dbar = dmean + 2 * mp;
  std::vector<double> hbase((size_t)mp, 0.0);
  if (mean_s)
    for (int64_t i = 0; i < M; ++i) hbase[i] = mean_s[i];
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(Xt, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, S));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(dbase, hbase.data(), (size_t)mp * sizeof(double), cudaMemcpyHostToDevice, S));
  const Spec &cs = m->has_cross ? m->cross_spec : m->spec;
  const double *Xtrain = m->has_cross ? m->Xk_cross : m->Xk;
  const double *Xtk = nullptr;
  DGP_TRY(scaled_points(ctx, S, cs, Xt, mp, dp, m->d, "Xts", &Xtk));

  // K* = k(X, X*) : no noise (AbstractGPRegressionModel.scala:279-286)
  CovParams cpx = make_cov_params(cs, false, m->d);
  const int gram = pick_gram(ctx, cs, cpx, dp, m->d, nmax);
  tick(ctx, 0, 0);
  NvtxRange r_predict("dgp:predict");
  {
    AssembleArgs a;
    a.X1 = Xtrain;  a.X2 = Xtk;  a.dp = dp;
    a.rows = m->N;  a.cols = M;  a.rows_p = np;  a.cols_p = mp;
    a.out = Ks;  a.ld = np;
    a.use_gram = gram;
    DGP_TRY(assemble(ctx, S, cpx, a));
    if (m->psi) DGP_TRY(add_basis_term(ctx, S, m->psi, np, m->psi_s, mp, m->basis_mp, Ks, np, np, mp, false));
  }
  tick(ctx, 0, 1);
  // mean = m(X*) + K*' alpha   (:463)
  DGP_TRY(gemv_t(ctx, S, Ks, np, np, mp, m->alpha, dbase, dmean));
  tick(ctx, 0, 2);
  double *C = nullptr;
  if (need_cov) {
    C = (double *)ctx->buf("pcov", (size_t)mp * mp * sizeof(double));
    if (!C) return DGP_ERR_OOM;
    // K** = k(X*, X*), no noise (:288-295); padding = identity
    AssembleArgs a;
    a.X1 = Xtk;  a.X2 = Xtk;  a.dp = dp;
    a.rows = M;  a.cols = M;  a.rows_p = mp;  a.cols_p = mp;
    a.out = C;  a.ld = mp;  a.symmetric = 1;  a.lower_only = 1;
    a.use_gram = gram;
    DGP_TRY(assemble(ctx, S, cpx, a));
    if (m->psi) DGP_TRY(add_basis_term(ctx, S, m->psi_s, mp, m->psi_s, mp, m->basis_mp, C, mp, mp, mp, true));
    // V' = K*' L^-T (blocked substitution from the right on tensor cores): K*' is assembled in its own right, test
    // points as rows, and Ks -- whose only use, the mean, is behind us on the stream -- receives V'
    DGP_BUF(KsT, double, ctx, "KsT", (size_t)np * mp * sizeof(double));
    {
      AssembleArgs at;
      at.X1 = Xtk;  at.X2 = Xtrain;  at.dp = dp;
      at.rows = M;  at.cols = m->N;  at.rows_p = mp;  at.cols_p = np;
      at.out = KsT;  at.ld = mp;
      at.use_gram = gram;
      DGP_TRY(assemble(ctx, S, cpx, at));
      if (m->psi) DGP_TRY(add_basis_term(ctx, S, m->psi_s, mp, m->psi, np, m->basis_mp, KsT, mp, mp, np, false));
    }
    double *Vt = Ks;
    DGP_TRY(trsm_right_lower_t(ctx, S, ctx->lookahead ? ctx->panel : S, m->L, np, m->inv, np, KsT, mp, Vt, mp, mp,
                               np >= 4096 ? 512 : 256));
    tick(ctx, 0, 3);
    GemmArgs s;  // Sigma* = K** - V' V  (lower tiles)   (:452-461)
    s.A = Vt;  s.lda = mp;  s.a_kmajor = 0;
    s.B = Vt;  s.ldb = mp;  s.b_kmajor = 0;
    s.C = C;  s.ldc = mp;
    s.M = (int)mp;  s.N = (int)mp;  s.K = (int)np;
    s.alpha = -1.0;  s.beta = 1.0;  s.lower_only = 1;
    DGP_TRY(gemm(ctx, S, s));
    DGP_TRY(mirror_lower(ctx, S, C, mp, mp));
    tick(ctx, 0, 4);
  }
  cudaEvent_t cov_ready = misc_event(ctx, EV_COV_READY);
  DGP_CUDA_OK(ctx, cudaEventRecord(cov_ready, S));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(mean, dmean, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, S));
  int32_t ret = DGP_OK;
  std::vector<double> hbar;
  if (lo) {
    // error bars: mean -/+ |sigma| * chol(Sigma*) . 1   (BlockedMultiVariateGaussian.scala:58-68); enqueued here,
    // finished on the host below, so that the second Cholesky runs underneath the copy-out of Sigma*
    DGP_BUF(Lp, double, ctx, "plpost", (size_t)mp * mp * sizeof(double));
    DGP_BUF(pinv, double, ctx, "pinv", (size_t)mp * TILE * sizeof(double));
    DGP_TRY(copy_block(ctx, S, C, mp, Lp, mp, mp, mp));
    DGP_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_info, 0, sizeof(int), S));
    DGP_TRY(potrf_blocked(ctx, Lp, mp, mp, pinv, ctx->d_info));
    DGP_TRY(lower_row_sums(ctx, S, Lp, mp, mp, dbar));
    hbar.resize((size_t)M);
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, S));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(hbar.data(), dbar, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, S));
  }
  if (cov) DGP_TRY(copy_out_2d(ctx, S, cov_ready, cov, ldcov, C, mp, M, M));
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  if (lo) {
    if (ctx->h_info[0] != 0) {
      for (int64_t i = 0; i < M; ++i) lo[i] = hi[i] = nan("");
      ret = ctx->fail(DGP_ERR_NOT_PD, "posterior covariance is not positive definite (pivot %d)", ctx->h_info[0] - 1);
    } else {
      const double mult = fabs(sigma);
      for (int64_t i = 0; i < M; ++i) {
        lo[i] = mean[i] - hbar[i] * mult;
        hi[i] = mean[i] + hbar[i] * mult;
      }
    }
  }
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  for (int i = 0; i < 16; ++i) ctx->timings[i] = 0.0;
  if (ctx->timing) {
    float ms;
    cudaEventElapsedTime(&ms, ctx->tev[0], ctx->tev[1]);  ctx->timings[5] = ms;
    if (need_cov) {
      cudaEventElapsedTime(&ms, ctx->tev[2], ctx->tev[3]);  ctx->timings[6] = ms;
      cudaEventElapsedTime(&ms, ctx->tev[3], ctx->tev[4]);  ctx->timings[7] = ms;
      cudaEventElapsedTime(&ms, ctx->tev[0], ctx->tev[4]);  ctx->timings[8] = ms;
    }
  }
  return ret;
}

// ---- kernel-matrix export ------------------------------------------------------------------------
int32_t dgp_kernel_matrix(dgp_ctx *ctx, const dgp_kernel_spec *spec, const double *X1, int64_t N1, const double *X2,
                          int64_t N2, int32_t d, int32_t x_layout, double *out, int64_t ld, int32_t add_noise) {
  if (!ctx || !spec || !X1 || !out || N1 <= 0 || d <= 0) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  const bool sym = X2 == nullptr;
  if (sym) N2 = N1;
  if (N2 <= 0 || ld < N1) return ctx->fail(DGP_ERR_SHAPE, "kernel_matrix: bad N2 / ld");
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, d, &sp));
  cudaStream_t S = ctx->stream;
  const int64_t n1p = round_up(N1, TILE), n2p = round_up(N2, TILE);
  const int dp = (int)round_up(d, 4);
  std::vector<double> h1, h2, ctr;
  column_means(X1, N1, d, x_layout, ctr);
  if (spec_uses_raw_points(sp)) ctr.assign(ctr.size(), 0.0);
  double nmax = pack_points(X1, N1, d, x_layout, n1p, dp, h1, ctr.data());
  DGP_BUF(A1, double, ctx, "km_x1", h1.size() * sizeof(double));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(A1, h1.data(), h1.size() * sizeof(double), cudaMemcpyHostToDevice, S));
  double *A2 = A1;
  if (!sym) {
    const double n2max = pack_points(X2, N2, d, x_layout, n2p, dp, h2, ctr.data());
    if (n2max > nmax || n2max != n2max) nmax = n2max;
    A2 = (double *)ctx->buf("km_x2", h2.size() * sizeof(double));
    if (!A2) return DGP_ERR_OOM;
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(A2, h2.data(), h2.size() * sizeof(double), cudaMemcpyHostToDevice, S));
  }
  const double *B1 = nullptr, *B2 = nullptr;
  DGP_TRY(scaled_points(ctx, S, sp, A1, n1p, dp, d, "km_x1s", &B1));
  if (sym) B2 = B1;
  else DGP_TRY(scaled_points(ctx, S, sp, A2, n2p, dp, d, "km_x2s", &B2));
  DGP_BUF(Kd, double, ctx, "km_out", (size_t)n1p * n2p * sizeof(double));
  CovParams cp = make_cov_params(sp, sym && add_noise, d);
  AssembleArgs a;
  a.X1 = B1;  a.X2 = B2;  a.dp = dp;
  a.rows = N1;  a.cols = N2;  a.rows_p = n1p;  a.cols_p = n2p;
  a.out = Kd;  a.ld = n1p;
  a.symmetric = sym ? 1 : 0;
  a.lower_only = sym ? 1 : 0;  // SVMKernel.scala:77-86 evaluates i >= j and mirrors: bitwise symmetric
  a.use_gram = pick_gram(ctx, sp, cp, dp, d, nmax);
  if (sym && add_noise && a.use_gram) a.nodup = points_stay_distinct(has_duplicate_rows(h1, N1, dp), sp, d);
  tick(ctx, 0, 0);
  DGP_TRY(assemble(ctx, S, cp, a));
  if (sym) DGP_TRY(mirror_lower(ctx, S, Kd, n1p, n1p));
  tick(ctx, 0, 1);
  DGP_TRY(copy_out_2d(ctx, S, nullptr, out, ld, Kd, n1p, N1, N2));
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  for (int i = 0; i < 16; ++i) ctx->timings[i] = 0.0;
  if (ctx->timing) {
    float ms;
    cudaEventElapsedTime(&ms, ctx->tev[0], ctx->tev[1]);
    ctx->timings[0] = ms;
  }
  return DGP_OK;
}

// ---- blocks of the training matrix from resident data (SVMKernel.buildPartitionedKernelMatrix, :181-219) ---------
int32_t dgp_data_kernel_block(dgp_ctx *ctx, const dgp_data *D, const dgp_kernel_spec *spec, int64_t row0, int64_t nrows,
                              int64_t col0, int64_t ncols, int32_t add_noise, double *out, int64_t ld) {
  if (!ctx || !D || !spec || !out) return DGP_ERR_ARG;
  if (row0 < 0 || col0 < 0 || nrows <= 0 || ncols <= 0 || row0 + nrows > D->N || col0 + ncols > D->N || ld < nrows)
    return ctx->fail(DGP_ERR_SHAPE, "kernel_block: rows [%lld, +%lld) x cols [%lld, +%lld) of an N = %lld set, ld = %lld",
                     (long long)row0, (long long)nrows, (long long)col0, (long long)ncols, (long long)D->N, (long long)ld);
  cudaSetDevice(ctx->device);
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, D->d, &sp));
  cudaStream_t S = ctx->stream;
  const int dp = D->dp;
  const int64_t rp = round_up(nrows, TILE), cpad = round_up(ncols, TILE);
  const bool diag = row0 == col0 && nrows == ncols;           // a diagonal block: symmetric, evaluated i >= j and mirrored
  const bool overlap = row0 < col0 + ncols && col0 < row0 + nrows;
  // the Dirac term fires wherever |x - y| == 0: on the diagonal, and anywhere when the set has duplicated points
  const bool noise = add_noise && (overlap || D->has_dup);
  const double *Xbase = spec_uses_raw_points(sp) ? D->Xraw : D->X;
  const double *Xk = nullptr;
  DGP_TRY(scaled_points(ctx, S, sp, Xbase, D->Np + TILE, dp, D->d, "kb_xs", &Xk));
  CovParams cp = make_cov_params(sp, noise, D->d);
  DGP_BUF(Kd, double, ctx, "kb_out", (size_t)rp * cpad * sizeof(double));
  AssembleArgs a;
  a.X1 = Xk + row0 * dp;  a.X2 = Xk + col0 * dp;  a.dp = dp;
  a.rows = nrows;  a.cols = ncols;  a.rows_p = rp;  a.cols_p = cpad;
  a.out = Kd;  a.ld = rp;
  a.symmetric = diag ? 1 : 0;
  a.lower_only = diag ? 1 : 0;
  // the Gram formulation needs 16-byte aligned operands (dp is a multiple of 4 doubles: any row offset is) and, with a
  // noise term, the guarantee that only i == j coincide: a diagonal block of a duplicate-free set
  a.use_gram = (!noise || (diag && !D->has_dup)) ? pick_gram(ctx, sp, cp, dp, D->d, D->max_sq_norm) : 0;
  a.nodup = diag ? points_stay_distinct(D->has_dup, sp, D->d) : 0;
  DGP_TRY(assemble(ctx, S, cp, a));
  if (diag) DGP_TRY(mirror_lower(ctx, S, Kd, rp, rp));
  DGP_TRY(copy_out_2d(ctx, S, nullptr, out, ld, Kd, rp, nrows, ncols));
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  return DGP_OK;
}

// ---- kernel matrix with its derivatives (SVMKernel.buildKernelGradMatrix / buildCrossKernelGradMatrix, :107-179) ---
int32_t dgp_kernel_grad_matrix(dgp_ctx *ctx, const dgp_kernel_spec *spec, const double *X1, int64_t N1, const double *X2,
                               int64_t N2, int32_t d, int32_t x_layout, int32_t add_noise, double *out) {
  if (!ctx || !spec || !X1 || !out || N1 <= 0 || d <= 0) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  const bool sym = X2 == nullptr;
  if (sym) N2 = N1;
  if (N2 <= 0) return ctx->fail(DGP_ERR_SHAPE, "kernel_grad_matrix: bad N2");
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, d, &sp));
  const int ng = dgp_spec_grad_len(spec);
  if (ng < 0) return ctx->fail(DGP_ERR_ARG, "kernel_grad_matrix: malformed spec");
  const size_t plane = (size_t)N1 * N2;
  // plane 0: the kernel matrix itself (key "kernel-matrix"), through the ordinary builder
  DGP_TRY(dgp_kernel_matrix(ctx, spec, X1, N1, sym ? nullptr : X2, N2, d, x_layout, out, N1, sym ? add_noise : 0));
  cudaStream_t S = ctx->stream;
  const int dp = (int)round_up(d, 4);
  const int64_t n1p = round_up(N1, TILE), n2p = round_up(N2, TILE);
  std::vector<double> h1, h2, ctr;
  column_means(X1, N1, d, x_layout, ctr);
  if (spec_uses_raw_points(sp)) ctr.assign(ctr.size(), 0.0);
  pack_points(X1, N1, d, x_layout, n1p, dp, h1, ctr.data());
  DGP_BUF(A1, double, ctx, "km_x1", h1.size() * sizeof(double));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(A1, h1.data(), h1.size() * sizeof(double), cudaMemcpyHostToDevice, S));
  double *A2 = A1;
  if (!sym) {
    pack_points(X2, N2, d, x_layout, n2p, dp, h2, ctr.data());
    A2 = (double *)ctx->buf("km_x2", h2.size() * sizeof(double));
    if (!A2) return DGP_ERR_OOM;
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(A2, h2.data(), h2.size() * sizeof(double), cudaMemcpyHostToDevice, S));
  }
  const double *B1 = nullptr, *B2 = nullptr;
  DGP_TRY(scaled_points(ctx, S, sp, A1, n1p, dp, d, "km_x1s", &B1));
  if (sym) B2 = B1;
  else DGP_TRY(scaled_points(ctx, S, sp, A2, n2p, dp, d, "km_x2s", &B2));
  // the derivative of the noise term exists only where the builder adds it (the symmetric training matrix)
  Spec spn = sp;
  if (!(sym && add_noise)) spn.noise = 0.0;
  CovParams cp = make_cov_params(spn, sym && add_noise, d);
  const int nsmax = cp.kind == DGP_KERNEL_COMPOSITE ? grad_partials_per_tile(cp, dp)
                                                    : std::max(grad_partials_per_tile(cp, dp), MAX_ARD_DIM + 2);
  // columns in chunks of at most ~64 MB of moments
  int64_t chunk = std::max<int64_t>(1, (int64_t)(64u << 20) / (int64_t)(sizeof(double) * nsmax * N1));
  if (chunk > N2) chunk = N2;
  DGP_BUF(mom, double, ctx, "kg_moments", (size_t)N1 * chunk * nsmax * sizeof(double));
  std::vector<double> hm((size_t)N1 * chunk * nsmax), g((size_t)ng + 1);
  for (int64_t c0 = 0; c0 < N2; c0 += chunk) {
    const int64_t nc = std::min(chunk, N2 - c0);
    int ns = 0;
    DGP_TRY(grad_pair_moments(ctx, S, cp, A1, B1, A2, B2, N1, c0, nc, dp, mom, &ns));
    DGP_CUDA_OK(ctx, cudaMemcpyAsync(hm.data(), mom, (size_t)N1 * nc * ns * sizeof(double), cudaMemcpyDeviceToHost, S));
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
    for (int64_t j = 0; j < nc; ++j)
      for (int64_t i = 0; i < N1; ++i) {
        grad_finish(spn, d, hm.data() + ((size_t)i + (size_t)j * N1) * ns, ns, g.data());
        for (int q = 0; q < ng; ++q) out[(size_t)(1 + q) * plane + (size_t)i + (size_t)(c0 + j) * N1] = g[q];
      }
  }
  return DGP_OK;
}

// ---- measurement hooks ---------------------------------------------------------------------------
int64_t dgp_outer_block(const dgp_ctx *ctx, int64_t n) {
  return ctx ? potrf_outer_block(ctx, round_up(n, TILE)) : 0;
}

int32_t dgp_launch_count(const dgp_ctx *ctx, uint64_t *out) {
  if (!ctx || !out) return DGP_ERR_ARG;
  *out = ctx->launches;
  return DGP_OK;
}

int32_t dgp_measure_peaks(dgp_ctx *ctx, double *out, int32_t n) {
  if (!ctx || !out) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  return measure_peaks(ctx, out, n);
}

int32_t dgp_time_phase(dgp_ctx *ctx, const dgp_data *D, const dgp_kernel_spec *spec, int32_t phase, int32_t reps,
                       double *avg_ms) {
  if (!ctx || !D || !spec || !avg_ms || reps <= 0) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  Spec sp;
  DGP_TRY(parse_spec(ctx, spec, D->d, &sp));
  cudaStream_t S = ctx->stream;
  const int64_t np = D->Np;
  DGP_BUF(K, double, ctx, "K", (size_t)np * np * sizeof(double));
  DGP_BUF(inv, double, ctx, "inv", (size_t)np * TILE * sizeof(double));
  const double *Xk = nullptr;
  DGP_TRY(scaled_points(ctx, S, sp, spec_uses_raw_points(sp) ? D->Xraw : D->X, np, D->dp, D->d, "Xs", &Xk));
  CovParams cp = make_cov_params(sp, true, D->d);
  AssembleArgs a;
  a.X1 = Xk;  a.X2 = Xk;  a.dp = D->dp;
  a.rows = D->N;  a.cols = D->N;  a.rows_p = np;  a.cols_p = np;
  a.out = K;  a.ld = np;  a.symmetric = 1;  a.lower_only = 1;
  a.use_gram = pick_gram(ctx, sp, cp, D->dp, D->d, D->max_sq_norm);
  a.nodup = points_stay_distinct(D->has_dup, sp, D->d);
  double total = 0.0;
  for (int r = 0; r < reps; ++r) {
    float ms = 0.f;
    if (phase == 0) {
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
      DGP_TRY(assemble(ctx, S, cp, a));
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
    } else if (phase == 1) {
      DGP_TRY(assemble(ctx, S, cp, a));
      DGP_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_info, 0, sizeof(int), S));
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
      DGP_TRY(potrf_blocked(ctx, K, np, np, inv, ctx->d_info));
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
    } else if (phase == 2) {
      // the dominant kernel in isolation: one full-size trailing update C -= P P^T with K = nb
      const int64_t nbw = potrf_outer_block(ctx, np);
      const int64_t w = nbw < np ? nbw : np;
      if (r == 0) DGP_TRY(assemble(ctx, S, cp, a));
      GemmArgs g;
      g.A = K + w;  g.lda = np;  g.B = K + w;  g.ldb = np;
      g.C = K + w + w * np;  g.ldc = np;
      g.M = (int)(np - w);  g.N = (int)(np - w);  g.K = (int)w;
      g.alpha = -1e-9;  g.beta = 1.0;  g.lower_only = 1;
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
      DGP_TRY(gemm(ctx, S, g));
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
    } else if (phase == 3) {
      // the largest single launch of the gradient path: K^-1 = U U^T over lower tiles with k >= row (N^3/3 flop, NT)
      DGP_BUF(W, double, ctx, "W", (size_t)np * np * sizeof(double));
      DGP_BUF(Kinv, double, ctx, "Kinv", (size_t)np * np * sizeof(double));
      if (r == 0) {
        DGP_TRY(assemble(ctx, S, cp, a));
        DGP_TRY(mirror_lower(ctx, S, K, np, np));
        DGP_TRY(copy_block(ctx, S, K, np, W, np, np, np));
      }
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
      DGP_TRY(lauum_upper_on(ctx, S, W, np, np, Kinv, np));
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
    } else if (phase == 4) {
      // the NN instantiation (TRMM / TRSM-type products of the triangular inverse and the prediction):
      // C = A B with M = N = K = min(np, 8192), full k-range (2 M^3 flop)
      const int64_t m = np < 8192 ? np : 8192;
      DGP_BUF(W, double, ctx, "W", (size_t)np * np * sizeof(double));
      if (r == 0) DGP_TRY(assemble(ctx, S, cp, a));
      GemmArgs g;
      g.A = K;  g.lda = np;  g.a_kmajor = 0;
      g.B = K;  g.ldb = np;  g.b_kmajor = 1;
      g.C = W;  g.ldc = np;
      g.M = (int)m;  g.N = (int)m;  g.K = (int)m;
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
      DGP_TRY(gemm(ctx, S, g));
      DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
    } else {
      return ctx->fail(DGP_ERR_ARG, "time_phase: unknown phase %d", phase);
    }
    DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
    DGP_CUDA_OK(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    total += ms;
  }
  *avg_ms = total / reps;
  return DGP_OK;
}

int32_t dgp_test_gemm(dgp_ctx *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, double alpha,
                      const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *C, int64_t ldc,
                      int32_t lower_only, double *ms_out) {
  if (!ctx || !A || !B || !C || M <= 0 || N <= 0 || K <= 0) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStream_t S = ctx->stream;
  const int64_t mp = round_up(M, TILE), npd = round_up(N, TILE), kp = round_up(K, TILE);
  // padded device copies, zero filled
  const int64_t a_rows = transA ? kp : mp, a_cols = transA ? mp : kp;
  const int64_t b_rows = transB ? npd : kp, b_cols = transB ? kp : npd;
  DGP_BUF(dA, double, ctx, "tg_a", (size_t)a_rows * a_cols * sizeof(double));
  DGP_BUF(dB, double, ctx, "tg_b", (size_t)b_rows * b_cols * sizeof(double));
  DGP_BUF(dC, double, ctx, "tg_c", (size_t)mp * npd * sizeof(double));
  DGP_CUDA_OK(ctx, cudaMemsetAsync(dA, 0, (size_t)a_rows * a_cols * sizeof(double), S));
  DGP_CUDA_OK(ctx, cudaMemsetAsync(dB, 0, (size_t)b_rows * b_cols * sizeof(double), S));
  DGP_CUDA_OK(ctx, cudaMemsetAsync(dC, 0, (size_t)mp * npd * sizeof(double), S));
  const int64_t ar = transA ? K : M, ac = transA ? M : K;
  const int64_t br = transB ? N : K, bc = transB ? K : N;
  DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(dA, a_rows * sizeof(double), A, lda * sizeof(double), ar * sizeof(double), ac,
                                     cudaMemcpyHostToDevice, S));
  DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(dB, b_rows * sizeof(double), B, ldb * sizeof(double), br * sizeof(double), bc,
                                     cudaMemcpyHostToDevice, S));
  DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(dC, mp * sizeof(double), C, ldc * sizeof(double), M * sizeof(double), N,
                                     cudaMemcpyHostToDevice, S));
  GemmArgs g;
  g.A = dA;  g.lda = a_rows;  g.a_kmajor = transA ? 1 : 0;
  g.B = dB;  g.ldb = b_rows;  g.b_kmajor = transB ? 0 : 1;
  g.C = dC;  g.ldc = mp;
  g.M = (int)mp;  g.N = (int)npd;  g.K = (int)kp;
  g.alpha = alpha;  g.beta = beta;
  g.lower_only = lower_only && mp == npd;
  DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
  DGP_TRY(gemm(ctx, S, g));
  DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
  DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(C, ldc * sizeof(double), dC, mp * sizeof(double), M * sizeof(double), N,
                                     cudaMemcpyDeviceToHost, S));
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  if (ms_out) *ms_out = ms;
  return DGP_OK;
}

int32_t dgp_test_potrf(dgp_ctx *ctx, int64_t n, double *A, int64_t lda, double *Linv, double *Ainv,
                       double *logdet_half, double *ms_out) {
  if (!ctx || !A || n <= 0 || lda < n) return DGP_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStream_t S = ctx->stream;
  const int64_t np = round_up(n, TILE);
  DGP_BUF(dA, double, ctx, "tp_a", (size_t)np * np * sizeof(double));
  DGP_BUF(inv, double, ctx, "tp_inv", (size_t)np * TILE * sizeof(double));
  DGP_BUF(d_scal, double, ctx, "scal", 4096 * 160 * sizeof(double));
  DGP_CUDA_OK(ctx, cudaMemsetAsync(dA, 0, (size_t)np * np * sizeof(double), S));
  DGP_CUDA_OK(ctx, cudaMemcpy2DAsync(dA, np * sizeof(double), A, lda * sizeof(double), n * sizeof(double), n,
                                     cudaMemcpyHostToDevice, S));
  pad_identity_kernel<<<(unsigned)((np * np + 255) / 256), 256, 0, S>>>(dA, np, n, np);
  DGP_CUDA_OK(ctx, cudaMemsetAsync(ctx->d_info, 0, sizeof(int), S));
  DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_a, S));
  DGP_TRY(potrf_blocked(ctx, dA, np, np, inv, ctx->d_info));
  DGP_CUDA_OK(ctx, cudaEventRecord(ctx->ev_b, S));
  DGP_TRY(diag_log_sum(ctx, S, dA, np, np, d_scal));
  double *dW = nullptr, *dKi = nullptr;
  if (Linv || Ainv) {
    dW = (double *)ctx->buf("tp_w", (size_t)np * np * sizeof(double));
    dKi = (double *)ctx->buf("tp_ki", (size_t)np * np * sizeof(double));
    if (!dW || !dKi) return DGP_ERR_OOM;
    DGP_CUDA_OK(ctx, cudaMemsetAsync(dW, 0, (size_t)np * np * sizeof(double), S));
    DGP_TRY(trtri_lower(ctx, dA, np, inv, np, dW, np, dKi, np));
    if (Ainv && ctx->inverse_transposed) {  // the gradient path's pair: U = L^-T in its own buffer, K^-1 = U U^T
      DGP_BUF(dU, double, ctx, "tp_u", (size_t)np * np * sizeof(double));
      DGP_TRY(trtri_lower_transposed_on(ctx, S, dA, np, inv, np, dU, np, dKi, np));
      DGP_TRY(lauum_upper_on(ctx, S, dU, np, np, dKi, np));
      DGP_TRY(mirror_lower(ctx, S, dKi, np, np));
    } else if (Ainv) {
      DGP_TRY(lauum_lower(ctx, dW, np, np, dKi, np));
      DGP_TRY(mirror_lower(ctx, S, dKi, np, np));
    }
  }
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_scal, d_scal, sizeof(double), cudaMemcpyDeviceToHost, S));
  DGP_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, S));
  DGP_CUDA_OK(ctx, cudaStreamSynchronize(S));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  if (ms_out) *ms_out = ms;
  if (ctx->h_info[0] != 0)
    return ctx->fail(DGP_ERR_NOT_PD, "Cholesky failed: pivot %d is not positive", ctx->h_info[0] - 1);
  // L: the strictly-upper tiles were never touched (they still hold the input); zero them on the host side
  DGP_CUDA_OK(ctx, cudaMemcpy2D(A, lda * sizeof(double), dA, np * sizeof(double), n * sizeof(double), n,
                                cudaMemcpyDeviceToHost));
  for (int64_t c = 0; c < n; ++c)
    for (int64_t r = 0; r < c; ++r) A[r + c * lda] = 0.0;
  if (Linv) {
    DGP_CUDA_OK(ctx, cudaMemcpy2D(Linv, n * sizeof(double), dW, np * sizeof(double), n * sizeof(double), n,
                                  cudaMemcpyDeviceToHost));
    for (int64_t c = 0; c < n; ++c)
      for (int64_t r = 0; r < c; ++r) Linv[r + c * n] = 0.0;
  }
  if (Ainv)
    DGP_CUDA_OK(ctx, cudaMemcpy2D(Ainv, n * sizeof(double), dKi, np * sizeof(double), n * sizeof(double), n,
                                  cudaMemcpyDeviceToHost));
  if (logdet_half) *logdet_half = ctx->h_scal[0];
  return DGP_OK;
}

}  // extern "C"
