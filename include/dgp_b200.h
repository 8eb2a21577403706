/*
 * dgp_b200.h -- C ABI of libdgp_b200.so: the B200-native (sm_100a) Gaussian-process hot path
 * that sits behind DynaML's dynaml-core API.
 *
 * DynaML has no FFI of its own; its extension mechanism is
 * subclass-and-override.  Each entry point below is what a JVM shim (JNI or JDK FFM, see
 * INTEGRATION.md) binds to replace one reference method.  Paths are relative to
 *   CORE = dynaml-core/src/main/scala/io/github/tailhq/dynaml
 *
 * Conventions
 *  - every function returns an int32 status (DGP_OK == 0); the message of the last failure on a
 *    context is available from dgp_last_error().
 *  - all matrices that cross the boundary are FP64, column-major, offset 0 (the layout of a
 *    Breeze DenseMatrix[Double] with isTranspose == false); `ld` is the major stride in elements.
 *  - point sets X are N x d.  x_layout == DGP_ROW_MAJOR means point i occupies
 *    X[i*d .. i*d+d-1] (a packed Seq[DenseVector[Double]]); DGP_COL_MAJOR means a Breeze N x d matrix.
 *  - the caller owns every host pointer; the library copies in/out before returning (synchronous).
 *  - a dgp_ctx is bound to ONE GPU and is NOT thread-safe: the reference model mutates kernel state in
 *    setState (CORE/models/gp/AbstractGPRegressionModel.scala:117-128).  Distinct contexts are independent:
 *    one context per host thread per GPU is a supported pattern (every entry point selects its context's
 *    device for the calling thread), and dgp_energy_batch_multi drives several contexts from one call.
 *  - there is no CPU fallback: without a usable sm_100 device dgp_ctx_create fails with DGP_ERR_CUDA.
 */
#ifndef DGP_B200_H
#define DGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dgp_ctx dgp_ctx;     /* one GPU, its streams and its workspace arena            */
typedef struct dgp_data dgp_data;   /* training inputs X, targets y, prior mean m(X), resident  */
typedef struct dgp_model dgp_model; /* a fitted model: L = chol(K), block inverses, alpha       */

/* ---- status codes ------------------------------------------------------------------------ */
enum {
  DGP_OK = 0,
  DGP_ERR_NOT_PD = 1,             /* Cholesky met a non-positive pivot: the reference's
                                     NotConvergedException (-> +Inf energy, NaN gradient,
                                     AbstractGPRegressionModel.scala:243-251, 531-533)          */
  DGP_ERR_UNSUPPORTED_KERNEL = 2,
  DGP_ERR_SHAPE = 3,
  DGP_ERR_CUDA = 4,
  DGP_ERR_NCCL = 5,
  DGP_ERR_OOM = 6,
  DGP_ERR_ARG = 7
};

enum { DGP_ROW_MAJOR = 0, DGP_COL_MAJOR = 1 };

/* ---- covariance functions ---------------------------------------------------------------- */
/* `hyp` is given in the order of the reference class's `hyper_parameters` list.               */
enum {
  DGP_KERNEL_SE = 0,       /* SEKernel            CORE/kernels/RBFKernel.scala:70-88
                              hyp = [bandwidth, amplitude]; k = a^2 exp(-|x-y|^2 / (2 l^2))     */
  DGP_KERNEL_RBF = 1,      /* RBFKernel           CORE/kernels/RBFKernel.scala:29-66
                              hyp = [bandwidth]                                                  */
  DGP_KERNEL_ARD_SE = 2,   /* MahalanobisKernel   CORE/kernels/RBFKernel.scala:101-146
                              hyp = [MahalanobisAmplitude, MahalanobisBandwidth_0 .. _{d-1}]     */
  DGP_KERNEL_MATERN = 3,   /* GenericMaternKernel CORE/kernels/GenericMaternKernel.scala:37-99
                              hyp = [p, l]; half-integer order nu = p + 1/2, 0 <= p <= 8         */
  DGP_KERNEL_PERIODIC = 4, /* PeriodicKernel      CORE/kernels/PeriodicKernel.scala:11-68
                              hyp = [sigma, lengthscale, frequency]; uses the L1 norm            */
  DGP_KERNEL_DIRAC = 5,    /* DiracKernel         CORE/kernels/DiracKernel.scala:30-60
                              hyp = [noiseLevel]; k = |s| * 1[|x-y|_2 == 0]                      */
  /* further stationary kernels of the same builder API, as assembly epilogues */
  DGP_KERNEL_CAUCHY = 6,   /* CauchyKernel        CORE/kernels/CauchyKernel.scala
                              hyp = [sigma]; k = 1 / (1 + |x-y|^2 / sigma^2)                     */
  DGP_KERNEL_LAPLACIAN = 7,/* LaplacianKernel     CORE/kernels/LaplacianKernel.scala
                              hyp = [beta]; k = exp(-|x-y|_1 / beta)   (L1 norm)                 */
  DGP_KERNEL_RATQUAD = 8,  /* RationalQuadraticKernel CORE/kernels/RationalQuadraticKernel.scala
                              hyp = [mu, l]; k = (1 + |x-y|^2/(mu l^2))^(-(d+mu)/2)              */
  DGP_KERNEL_TSTUDENT = 9, /* TStudentKernel      CORE/kernels/TStudentKernel.scala
                              hyp = [d]; k = 1 / (1 + |x-y|^d)                                   */
  /* Kernel algebra of LocalScalarKernel (`+`, `*`, `* c`; CORE/kernels/LocalScalarKernel.scala:53-91,
     AdditiveCovariance.scala:27-93, MultiplicativeCovariance.scala:11-90) in sum-of-products normal
     form  k = sum_t coef_t * prod_f k_f ^ e_tf  over F <= DGP_MAX_FACTORS base kernels (any kind
     above) and T <= DGP_MAX_TERMS terms.  `hyp` is then a flat program of n_hyp doubles:
        [ F, T,
          F x { kind, grad_mode, n_f, hyp_0 .. hyp_{n_f - 1} },        base kernels, reference order
          T x { coef_t, e_t0 .. e_t,F-1 },                            small non-negative integers
          optionally F x { start_f, count_f } ]                       tensor kernels, see below
     gradEnergy returns the derivatives of the base kernels' hyper-parameters factor by factor,
     then the noise level: dgp_spec_grad_len() values.
     The optional trailer restricts base kernel f to the coordinates [start_f, start_f + count_f) of the input
     vector (count 0 = to the end): kernels on tuple inputs, TensorCombinationKernel / KroneckerProductKernel
     (CORE/kernels/TensorCombinationKernel.scala:12-97) over the concatenated vectors; an ARD-SE factor then has
     count_f bandwidths.  ARD-SE and sliced factors need d <= 32 and at most four distinct metrics.        */
  DGP_KERNEL_COMPOSITE = 10,
  /* The two non-stationary kernels of SURVEY 8(f) rank 1.  They read the original (uncentred) points and run
     on the composite evaluation path (also usable as base kernels of a composite).                          */
  DGP_KERNEL_POLYNOMIAL = 11, /* PolynomialKernel  CORE/kernels/PolynomialKernel.scala:9-45
                              hyp = [degree, offset]; k = (x.y + offset)^int(degree); the reference
                              defines no gradient (LocalSVMKernel.scala:37-41 stub): derivatives are 0 */
  DGP_KERNEL_FBM = 12      /* FBMKernel           CORE/kernels/FBMKernel.scala:70-122
                              hyp = [hurst]; k = (|x|^2H + |y|^2H - |x-y|^2H) / 2.  DGP_GRAD_REFERENCE
                              reproduces the reference derivative a^H ln a + b^H ln b - c^H ln c with
                              c = |x-y|^2 (no 1/2; NaN wherever c == 0, hence NaN for every training set);
                              DGP_GRAD_EXACT gives the true one (1/2, and 0 ln 0 = 0)                  */
};

#define DGP_MAX_FACTORS 32
#define DGP_MAX_TERMS 64

/* Gradient conventions for hyper-parameters whose reference derivative is defective. */
enum {
  DGP_GRAD_REFERENCE = 0,  /* bit-for-bit the reference formulas, including
                              MahalanobisKernel's k*|x-y|^2/b_i^3 for every i (RBFKernel.scala:142-143) */
  DGP_GRAD_EXACT = 1       /* the mathematically exact dk/db_i = k*(x_i-y_i)^2/b_i^3             */
};

#define DGP_MAX_HYP 130    /* 1 amplitude + up to 128 ARD bandwidths + slack */

typedef struct dgp_kernel_spec {
  int32_t kind;            /* DGP_KERNEL_*                                                      */
  int32_t n_hyp;           /* entries of hyp                                                    */
  int32_t grad_mode;       /* DGP_GRAD_*                                                        */
  int32_t reserved;
  const double *hyp;       /* covariance hyper-parameters, reference order                      */
  double noise;            /* DiracKernel noiseLevel of the noise model (signed; |.| is added
                              wherever |x-y|_2 == 0, DiracKernel.scala:43-47)                   */
} dgp_kernel_spec;

/* Number of values dgp_energy_grad writes for this spec (n_hyp + 1, or for a composite the sum of
   its base kernels' hyper-parameter counts + 1); negative DGP_ERR_* on a malformed spec.       */
int32_t dgp_spec_grad_len(const dgp_kernel_spec *spec);

/* ---- context ------------------------------------------------------------------------------ */
int32_t dgp_ctx_create(int32_t device, dgp_ctx **out);
int32_t dgp_ctx_destroy(dgp_ctx *ctx);
const char *dgp_last_error(const dgp_ctx *ctx);     /* NUL-terminated, owned by ctx             */
/* Return the context's grow-only workspace (kernel matrix, inverse, panels, tile tables) to the driver.
   Data and model handles stay valid; the next evaluation allocates again.  The analogue of the reference's
   unpersist (AbstractGPRegressionModel.scala:418-422) for the library's own scratch.                  */
int32_t dgp_ctx_trim(dgp_ctx *ctx);
/* out[0] CUDA device ordinal, out[1] SM count, out[2] total / out[3] free device memory in bytes; n >= 4. */
int32_t dgp_ctx_info(dgp_ctx *ctx, int64_t *out, int32_t n);
const char *dgp_version(void);
/* Tunables: "nb" (outer Cholesky block, multiple of 128; 0 = auto, the default: chosen per outer block from the size
   of what is left to factor), "lookahead"
   (0/1), "timing" (0/1), "batch_streams" (1..4 candidates of dgp_energy_batch in flight; 0 = by size, the default: 4 up to N = 12288, else 3),
   "assembly" (0 = Gram/DMMA kernel when its error bound allows, 1 = exact differences, 2 = force Gram).
   GEMM kernel selection (results are bit-identical for every choice): "gemm_variant" (4 = bulk-copy /
   mbarrier kernel where "gemm_v4_mask" allows, the default; 2 = cp.async kernel everywhere), "gemm_v4_mask" and
   "gemm_v4_bk32_mask" (bit 2*a_kmajor + b_kmajor per operand layout: layouts on the bulk-copy kernel, and
   those of them using 32-wide K-slices; defaults 3 and 2), "gemm_bk" (16/32, cp.async kernel),
   "gemm_handover" (stage hand-over of the bulk-copy kernel), "gemm_no_prefetch" (0/1, testing: no L2 prefetch of the C tile),
   "inverse_transposed" (1 = K^-1 as U U^T from U = L^-T, the default; 0 = W^T W from W = L^-1),
   "graphs" (1 = the blocked Cholesky of energy / energy_grad / energy_batch is replayed from a CUDA graph after
   its second use at a given size; 0 = launched kernel by kernel, the default -- the graph replay measured 3-12 %
   slower on one GPU; same results either way),
   "leaf_variant" (1 = tensor-pipe 128-block leaf, the default; 0 = scalar leaf),
   "trsv_chained" (1 = each triangular solve is one launch whose CTAs hand the solved segments over through
   flags, the default; 0 = one launch per 128-row block; same results),
   "assembly_store" (lean SE assembly: 0 = 16-byte stores from registers, the default; 1 = staged chunk + bulk
   stores, measured slower),
   "grad_lean" (gradient trace pass: 1 = from 4096 points up the strictly-lower tiles take the lean Gram pass under
   the conditions of the lean assembly -- Gram accuracy gate, pairwise distinct points; SE / RBF, ARD-SE in the
   reference gradient mode, Matern orders 1-3 -- the default; 0 = exact differences everywhere; the two agree to
   ~1e-13 relative).                                                                                        */
int32_t dgp_ctx_set_option(dgp_ctx *ctx, const char *key, int64_t value);
/* Phase timings (ms, CUDA events) of the last energy/energy_grad/fit/predict call:
   [0] assemble [1] potrf [2] solves+logdet [3] inverse (trtri+lauum) [4] trace pass
   [5] cross assemble [6] trsm [7] syrk [8] total; after dgp_energy_batch [9] is the device time of the whole
   batch.  Unused slots are 0.                                                                    */
int32_t dgp_get_timings(const dgp_ctx *ctx, double *ms, int32_t n);

/* ---- training data (replaces the Seq[(DenseVector[Double], Double)] the model holds,
        CORE/models/gp/GPRegression.scala:51-60; packed once, resident in HBM) ---------------- */
int32_t dgp_data_create(dgp_ctx *ctx, const double *X, int64_t N, int32_t d, int32_t x_layout,
                        const double *y, const double *mean /* m(X), NULL => 0 */,
                        dgp_data **out);
int32_t dgp_data_destroy(dgp_ctx *ctx, dgp_data *data);

/* ---- LocalScalarKernel.buildKernelMatrix / buildCrossKernelMatrix
        (CORE/kernels/LocalScalarKernel.scala:154-158 -> SVMKernel.scala:71-105) ---------------
   X2 == NULL: symmetric N1 x N1 matrix of X1 (both triangles written); add_noise != 0 adds the
   Dirac term of spec->noise.  Otherwise the N1 x N2 cross matrix (never any noise).             */
int32_t dgp_kernel_matrix(dgp_ctx *ctx, const dgp_kernel_spec *spec,
                          const double *X1, int64_t N1, const double *X2, int64_t N2,
                          int32_t d, int32_t x_layout, double *out, int64_t ld, int32_t add_noise);

/* ---- SVMKernel.buildPartitionedKernelMatrix (CORE/kernels/SVMKernel.scala:181-219; LocalScalarKernel.buildBlockedKernelMatrix,
        LocalScalarKernel.scala:160-161): the training matrix block by block, straight from the resident points --------
   out (nrows x ncols, column-major, leading dimension ld) = K[row0 : row0 + nrows, col0 : col0 + ncols] of the kernel
   matrix of the points held by `data`; add_noise != 0 adds the Dirac term of spec->noise wherever two points coincide
   (the diagonal; and any duplicates).  A block with equal row and column ranges is evaluated for i >= j and mirrored,
   as buildSVMKernelMatrix does (bitwise symmetric); any other block as crossKernelMatrix does.  No re-upload per block.
   The reference's PartitionedPSDMatrix is the list of such blocks for block row >= block column with
   num_elements_per_row_block = num_elements_per_col_block = blockSize.                                            */
int32_t dgp_data_kernel_block(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, int64_t row0,
                              int64_t nrows, int64_t col0, int64_t ncols, int32_t add_noise, double *out, int64_t ld);

/* ---- SVMKernel.buildKernelGradMatrix / buildCrossKernelGradMatrix (CORE/kernels/SVMKernel.scala:107-179) -----------
   The kernel matrix and its derivative with respect to every hyper-parameter, per entry.  `out` receives
   1 + dgp_spec_grad_len(spec) planes of N1 x N2 (column-major, contiguous): plane 0 the kernel matrix (the
   reference's key "kernel-matrix"), planes 1.. the derivatives in the order of dgp_energy_grad (the spec's
   hyper-parameters, then the noise level; NaN where the reference defines none).  X2 == NULL: the symmetric matrix
   of X1, with the Dirac term and its derivative when add_noise != 0.  Meant for inspection and parity at moderate
   N (the derivatives pass through the host); gradEnergy itself never materialises them.                          */
int32_t dgp_kernel_grad_matrix(dgp_ctx *ctx, const dgp_kernel_spec *spec, const double *X1, int64_t N1,
                               const double *X2, int64_t N2, int32_t d, int32_t x_layout, int32_t add_noise, double *out);

/* ---- AbstractGPRegressionModel.energy (AbstractGPRegressionModel.scala:148-189, 516-535) ----
   E = 0.5 * ( r' K^-1 r + sum_i log L_ii + n log 2 pi ),  r = y - m(X)  (the reference's value,
   whose log-det term is half the textbook one).  Non-PD K: *e = +Inf, returns DGP_ERR_NOT_PD.   */
int32_t dgp_energy(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, double *e);

/* The two data-dependent terms every likelihood on this path is built from:
     *quad = r' K^-1 r,   *logdet_half = sum_i log L_ii.
   GP: E = 0.5 (quad + logdet_half + n log 2 pi); Student-t process (SURVEY 8f rank 2,
   CORE/models/stp/AbstractSTPRegressionModel.scala:253-283,
   CORE/probability/distributions/BlockedMultivariateStudentsT.scala:47-61): the host combines the
   same two numbers with the degrees of freedom.  Non-PD K: DGP_ERR_NOT_PD.                       */
int32_t dgp_energy_terms(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, double *quad,
                         double *logdet_half);

/* ---- general noise models ----------------------------------------------------------------------
   AbstractGPRegressionModel takes any LocalScalarKernel as noise model: the training matrix is built from
   covariance + noiseModel, the cross and test matrices from the covariance alone (:269-295).  A DiracKernel
   noise is the `noise` field of the spec.  For any other noise kernel fit the model with the composite
   covariance + noise (noise = 0) and hand the covariance-only spec to dgp_model_set_cross_spec: dgp_predict
   and dgp_model_cross_apply then use it for k(X, X*) and k(X*, X*).                                     */
int32_t dgp_model_set_cross_spec(dgp_ctx *ctx, dgp_model *model, const dgp_kernel_spec *spec);

/* ---- building blocks for sibling models that share the Cholesky / solve core (SURVEY 8f rank 2:
        ESGPModel.solve, CORE/models/sgp/ESGPModel.scala:349-389; the `\\` solvers of
        CORE/algebra/PartitionedMatrixSolvers.scala) ----------------------------------------------
   dgp_model_solve:       out = (K + noise)^-1 B for k right-hand sides (B, out column-major n x k) using the
                          factor kept by dgp_fit  (Lmat.t \\ (Lmat \\ b)).
   dgp_model_cross_apply: out = k(X, X*)' V for k vectors (V column-major n x k, out n_test x k): crossKernel.t * v.
   Both copy in / out of host memory and synchronise.                                                */
int32_t dgp_model_solve(dgp_ctx *ctx, const dgp_model *model, const double *b, int32_t k, double *out);
int32_t dgp_model_cross_apply(dgp_ctx *ctx, const dgp_model *model, const double *xs, int64_t n_test,
                              int32_t x_layout, const double *v, int32_t k, double *out);

/* ---- explicit basis functions (GPBasisFuncRegressionModel,
        CORE/models/gp/GPBasisFuncRegressionModel.scala:70-109) ------------------------------------
   The model adds feature_map_cov(x, y) = psi(x).psi(y), psi(x) = chol(covB) * basisFunc(x), to the training,
   cross and test kernel matrices (getTrainKernelMatrix / getCrossKernelMatrix / getTestKernelMatrix overrides).
   The host evaluates basisFunc (an arbitrary user function) and hands over Psi, row-major n x m; the device
   adds Psi Psi' as a rank-m DMMA update after the assembly.  dgp_data_set_basis(m = 0) removes it.
   A model fitted from such data needs the test points' features before every dgp_predict.            */
int32_t dgp_data_set_basis(dgp_ctx *ctx, dgp_data *data, const double *psi, int32_t m);
int32_t dgp_model_set_test_basis(dgp_ctx *ctx, dgp_model *model, const double *psi_s, int64_t n_test, int32_t m);

/* ---- gradEnergy (AbstractGPRegressionModel.scala:196-267) -----------------------------------
   g[i] = tr( (alpha alpha' - K^-1) dK/dtheta_i ), i over spec->hyp in order, then g[n_hyp] for
   the noise level: n_hyp + 1 values (entries with no reference derivative, e.g. Matern "p",
   are NaN).  No -1/2 factor, exactly as the reference.  Non-PD: all NaN + DGP_ERR_NOT_PD.
   `e` may be NULL.                                                                               */
int32_t dgp_energy_grad(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec,
                        double *e, double *g);

/* ---- the candidate loop of ModelTuner.getEnergyLandscape / AbstractCSA
        (CORE/optimization/ModelTuner.scala:131-153, AbstractCSA.scala:214-259) ----------------
   Evaluates n candidates; status[i] is per candidate so one non-PD candidate yields +Inf
   without aborting the batch.  Returns DGP_OK unless the batch itself could not run.           */
int32_t dgp_energy_batch(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *specs,
                         int32_t n, double *e, int32_t *status);

/* The same loop over several GPUs from ONE process (DynaML is a single JVM): ctxs[c] is a context on its own
   device and data[c] the same training set created on it; candidate i runs on context i mod n_ctx, the shards
   concurrently (one internal host thread per context), and only the energies travel.  Results are bit-identical
   to dgp_energy_batch on one context.  The caller must not use any of the listed contexts from another thread
   during the call.  Errors of context c > 0 are reported through ctxs[0].                               */
int32_t dgp_energy_batch_multi(dgp_ctx *const *ctxs, const dgp_data *const *data, int32_t n_ctx,
                               const dgp_kernel_spec *specs, int32_t n, double *e, int32_t *status);

/* ---- persist + predictiveDistribution + predictionWithErrorBars
        (AbstractGPRegressionModel.scala:304-413, 434-466;
         CORE/probability/distributions/BlockedMultiVariateGaussian.scala:58-68) ---------------
   dgp_fit keeps L, the diagonal-block inverses and alpha resident (the reference caches K and
   refactors on every call; same observable results).
   dgp_predict:  mean[M]  = m(X*) + K*' alpha             (mean_s = m(X*), NULL => 0)
                 cov[MxM] = K** - (L^-1 K*)' (L^-1 K*)    (both triangles; NULL to skip)
                 lo/hi[M] = mean -/+ |sigma| * (chol(cov) . 1)   (NULL to skip; the reference's
                            row-sum error bars, not sqrt(diag))
   A posterior covariance that is not numerically PD makes lo/hi fail with DGP_ERR_NOT_PD
   (the reference throws from bcholesky there); mean and cov are still valid.                    */
int32_t dgp_fit(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec, dgp_model **out);
int32_t dgp_model_destroy(dgp_ctx *ctx, dgp_model *model);
int32_t dgp_model_energy(dgp_ctx *ctx, const dgp_model *model, double *e);
int32_t dgp_model_terms(dgp_ctx *ctx, const dgp_model *model, double *quad, double *logdet_half);
int32_t dgp_predict(dgp_ctx *ctx, const dgp_model *model, const double *Xs, int64_t M,
                    int32_t x_layout, const double *mean_s, double *mean, double *cov,
                    int64_t ldcov, double *lo, double *hi, double sigma);

/* ---- multi-GPU: 2D block-cyclic Cholesky + NLL over NCCL (one process per GPU) --------------
   Replaces nothing in the reference (its only distributed backend is Spark RDDs and is not on
   the GP path); it is the N > one-GPU route to the same `energy` value.
   Bootstrap: rank 0 calls dgp_dist_unique_id, the host side broadcasts the 128 bytes
   (torch.distributed / any side channel), every rank calls dgp_dist_init.                      */
int32_t dgp_dist_unique_id(void *id128);
int32_t dgp_dist_init(dgp_ctx *ctx, const void *id128, int32_t rank, int32_t world,
                      int32_t grid_rows, int32_t grid_cols);
int32_t dgp_dist_finalize(dgp_ctx *ctx);
/* out[0] ranks of the library's NCCL communicator (0: none, world == 1), out[1] this context's rank in it,
   out[2] / out[3] process grid rows / columns, out[4] 1 if the grid is emulated on one GPU, out[5] NCCL version
   code; n >= 6.                                                                                           */
int32_t dgp_dist_info(dgp_ctx *ctx, int32_t *out, int32_t n);
/* Host-only layout query (no GPU needed): out[0] tiles per dimension, out[1] / out[2] local tile rows /
   columns of `rank`, out[3] lower-triangle tiles it owns, out[4] bytes of its local matrix.           */
int32_t dgp_dist_plan(int64_t n, int32_t tile, int32_t grid_rows, int32_t grid_cols, int32_t rank, int64_t *out);
/* X, y, mean are replicated on every rank (N*d doubles); K is never materialised on the host.
   tile = block-cyclic tile edge (multiple of 256, at most 2048).  With world == 1 and
   grid_rows*grid_cols > 1 in dgp_dist_init, all ranks of the grid are emulated on this one GPU
   (same code path, broadcasts become device copies) -- used by the parity tests.  Every rank gets the same *e.
   chol_ms (nullable) receives this rank's device time of the factorisation alone.              */
int32_t dgp_dist_energy(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec,
                        int32_t tile, double *e, double *chol_ms);

/* ---- measurement hooks (bench.py / tests; not part of the drop-in surface) ------------------ */
/* out[0] DMMA FP64 TFLOP/s, out[1] DFMA FP64 TFLOP/s, out[2] HBM write GB/s, out[3] HBM copy GB/s
   (read+write bytes), out[4] HBM read GB/s, each the best of a few timed launches; out[5] DMMA TFLOP/s
   sustained over ~1 s of back-to-back launches; out[6], out[7] DMMA and DFMA TFLOP/s when interleaved
   in one kernel (do the two FP64 pipes add up?); out[8] DMMA TFLOP/s with the register pattern and
   occupancy of the GEMM kernel (4x4 accumulators, fresh fragments every step, 16 warps per SM).              */
int32_t dgp_measure_peaks(dgp_ctx *ctx, double *out, int32_t n);
/* Number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
int32_t dgp_launch_count(const dgp_ctx *ctx, uint64_t *out);
/* Outer block (trailing-update K) the blocked Cholesky uses for an n x n matrix on this context. */
int64_t dgp_outer_block(const dgp_ctx *ctx, int64_t n);
/* Time `reps` launches of one phase on device-resident data: returns average ms per launch.
   phase: 0 assemble (lower tiles of K+noise), 1 potrf of the assembled K, 2 trailing-update GEMM
   at full size (M = N - nb rows, K = nb = dgp_outer_block(N): the NT launch the Cholesky repeats), 3 the
   K^-1 = U U^T launch of the gradient path (NT, k >= row, N^3/3 flop: the largest single launch of a step),
   4 an NN product with M = N = K = min(N, 8192) (2 M^3 flop).  Used for the roofline of the dominant kernels. */
int32_t dgp_time_phase(dgp_ctx *ctx, const dgp_data *data, const dgp_kernel_spec *spec,
                       int32_t phase, int32_t reps, double *avg_ms);
/* Plain column-major FP64 GEMM / Cholesky on host buffers through the device kernels, for the
   parity tests of the building blocks.  op flags: 0 = as stored (M x K / K x N), 1 = transposed
   storage.  lower_only != 0 computes only tiles on or below the diagonal.                       */
int32_t dgp_test_gemm(dgp_ctx *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                      double alpha, const double *A, int64_t lda, const double *B, int64_t ldb,
                      double beta, double *C, int64_t ldc, int32_t lower_only, double *ms);
/* In: symmetric PD A (n x n, lower triangle read).  Out: L in the lower triangle of A (upper
   zeroed), and, if Linv != NULL, L^-1 (lower, upper zeroed); if Ainv != NULL, A^-1 (both
   triangles).  logdet_half = sum_i log L_ii.                                                    */
int32_t dgp_test_potrf(dgp_ctx *ctx, int64_t n, double *A, int64_t lda, double *Linv, double *Ainv,
                       double *logdet_half, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* DGP_B200_H */
