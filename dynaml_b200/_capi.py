"""ctypes binding of libdgp_b200.so (include/dgp_b200.h).

This is the Python twin of the JNI/FFM stub shown in INTEGRATION.md: plain pointers and sizes,
no torch types.  There is deliberately NO fallback: if the CUDA library is missing or no sm_100
device is usable, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libdgp_b200.so"

DGP_OK, DGP_ERR_NOT_PD, DGP_ERR_UNSUPPORTED_KERNEL, DGP_ERR_SHAPE = 0, 1, 2, 3
DGP_ERR_CUDA, DGP_ERR_NCCL, DGP_ERR_OOM, DGP_ERR_ARG = 4, 5, 6, 7
DGP_ROW_MAJOR, DGP_COL_MAJOR = 0, 1
KERNEL_SE, KERNEL_RBF, KERNEL_ARD_SE, KERNEL_MATERN, KERNEL_PERIODIC, KERNEL_DIRAC = range(6)
KERNEL_CAUCHY, KERNEL_LAPLACIAN, KERNEL_RATQUAD, KERNEL_TSTUDENT = 6, 7, 8, 9
GRAD_REFERENCE, GRAD_EXACT = 0, 1

_STATUS_NAMES = {
    0: "DGP_OK", 1: "DGP_ERR_NOT_PD", 2: "DGP_ERR_UNSUPPORTED_KERNEL", 3: "DGP_ERR_SHAPE",
    4: "DGP_ERR_CUDA", 5: "DGP_ERR_NCCL", 6: "DGP_ERR_OOM", 7: "DGP_ERR_ARG",
}

# every symbol include/dgp_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "dgp_ctx_create", "dgp_ctx_destroy", "dgp_last_error", "dgp_version", "dgp_ctx_set_option",
    "dgp_get_timings", "dgp_data_create", "dgp_data_destroy", "dgp_kernel_matrix", "dgp_energy", "dgp_energy_terms",
    "dgp_energy_grad", "dgp_energy_batch", "dgp_fit", "dgp_model_destroy", "dgp_model_energy", "dgp_model_terms",
    "dgp_predict", "dgp_dist_unique_id", "dgp_dist_init", "dgp_dist_finalize", "dgp_dist_plan", "dgp_dist_energy",
    "dgp_measure_peaks", "dgp_launch_count", "dgp_outer_block", "dgp_time_phase", "dgp_test_gemm", "dgp_test_potrf",
    "dgp_spec_grad_len", "dgp_data_set_basis", "dgp_model_set_test_basis", "dgp_model_solve", "dgp_model_cross_apply",
    "dgp_model_set_cross_spec", "dgp_ctx_trim", "dgp_dist_info", "dgp_ctx_info", "dgp_energy_batch_multi", "dgp_data_kernel_block", "dgp_kernel_grad_matrix",
]


class DgpError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{_STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class NotPositiveDefinite(DgpError):
    """The reference's breeze.linalg.NotConvergedException on this path."""


class KernelSpec(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_hyp", C.c_int32), ("grad_mode", C.c_int32), ("reserved", C.c_int32),
        ("hyp", C.POINTER(C.c_double)), ("noise", C.c_double),
    ]


_lib = None
_dp = C.POINTER(C.c_double)


def load_library() -> C.CDLL:
    """dlopen the in-tree library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dynaml_b200.build` "
            "(there is no CPU fallback for the GP hot path)")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sp = C.POINTER(KernelSpec)
    sig = {
        "dgp_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "dgp_ctx_destroy": (i32, [vp]),
        "dgp_last_error": (C.c_char_p, [vp]),
        "dgp_version": (C.c_char_p, []),
        "dgp_ctx_set_option": (i32, [vp, C.c_char_p, i64]),
        "dgp_get_timings": (i32, [vp, _dp, i32]),
        "dgp_data_create": (i32, [vp, _dp, i64, i32, i32, _dp, _dp, C.POINTER(vp)]),
        "dgp_data_destroy": (i32, [vp, vp]),
        "dgp_kernel_matrix": (i32, [vp, sp, _dp, i64, _dp, i64, i32, i32, _dp, i64, i32]),
        "dgp_data_kernel_block": (i32, [vp, vp, sp, i64, i64, i64, i64, i32, _dp, i64]),
        "dgp_kernel_grad_matrix": (i32, [vp, sp, _dp, i64, _dp, i64, i32, i32, i32, _dp]),
        "dgp_energy": (i32, [vp, vp, sp, _dp]),
        "dgp_energy_terms": (i32, [vp, vp, sp, _dp, _dp]),
        "dgp_model_terms": (i32, [vp, vp, _dp, _dp]),
        "dgp_energy_grad": (i32, [vp, vp, sp, _dp, _dp]),
        "dgp_energy_batch": (i32, [vp, vp, sp, i32, _dp, C.POINTER(i32)]),
        "dgp_energy_batch_multi": (i32, [C.POINTER(vp), C.POINTER(vp), i32, sp, i32, _dp, C.POINTER(i32)]),
        "dgp_fit": (i32, [vp, vp, sp, C.POINTER(vp)]),
        "dgp_model_destroy": (i32, [vp, vp]),
        "dgp_model_energy": (i32, [vp, vp, _dp]),
        "dgp_predict": (i32, [vp, vp, _dp, i64, i32, _dp, _dp, _dp, i64, _dp, _dp, f64]),
        "dgp_dist_unique_id": (i32, [vp]),
        "dgp_dist_init": (i32, [vp, vp, i32, i32, i32, i32]),
        "dgp_dist_finalize": (i32, [vp]),
        "dgp_dist_info": (i32, [vp, C.POINTER(i32), i32]),
        "dgp_ctx_trim": (i32, [vp]),
        "dgp_ctx_info": (i32, [vp, C.POINTER(i64), i32]),
        "dgp_dist_energy": (i32, [vp, vp, sp, i32, _dp, _dp]),
        "dgp_dist_plan": (i32, [i64, i32, i32, i32, i32, C.POINTER(i64)]),
        "dgp_measure_peaks": (i32, [vp, _dp, i32]),
        "dgp_launch_count": (i32, [vp, C.POINTER(C.c_uint64)]),
        "dgp_outer_block": (i64, [vp, i64]),
        "dgp_time_phase": (i32, [vp, vp, sp, i32, i32, _dp]),
        "dgp_test_gemm": (i32, [vp, i32, i32, i64, i64, i64, f64, _dp, i64, _dp, i64, f64, _dp, i64, i32, _dp]),
        "dgp_test_potrf": (i32, [vp, i64, _dp, i64, _dp, _dp, _dp, _dp]),
        "dgp_spec_grad_len": (i32, [sp]),
        "dgp_data_set_basis": (i32, [vp, vp, _dp, i32]),
        "dgp_model_set_test_basis": (i32, [vp, vp, _dp, i64, i32]),
        "dgp_model_solve": (i32, [vp, vp, _dp, i32, _dp]),
        "dgp_model_set_cross_spec": (i32, [vp, vp, sp]),
        "dgp_model_cross_apply": (i32, [vp, vp, _dp, i64, i32, _dp, i32, _dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a, order="C"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "C" if order == "C" else "F"])


def make_spec(kind: int, hyp, noise: float = 0.0, grad_mode: int = GRAD_REFERENCE):
    """Returns (KernelSpec, keepalive) -- keep the second alive while the spec is in use."""
    h = np.ascontiguousarray(np.asarray(hyp, dtype=np.float64).ravel())
    s = KernelSpec(kind, h.size, grad_mode, 0, h.ctypes.data_as(_dp), float(noise))
    return s, h


KERNEL_COMPOSITE, KERNEL_POLYNOMIAL, KERNEL_FBM = 10, 11, 12
MAX_FACTORS, MAX_TERMS = 32, 64


def make_composite_spec(factors, terms, noise: float = 0.0, ranges=None):
    """DGP_KERNEL_COMPOSITE: k = sum_t coef_t prod_f k_f^e_tf.  `factors` = [(kind, grad_mode, hyp), ...];
    `terms` = [(coef, [e_0 .. e_{F-1}]), ...]; `ranges` (optional) = per factor the (start, count) slice of the
    input vector it sees, count 0 = to the end.  Returns (KernelSpec, keepalive) like make_spec."""
    if not (1 <= len(factors) <= MAX_FACTORS and 1 <= len(terms) <= MAX_TERMS):
        raise ValueError(f"a composite kernel takes 1..{MAX_FACTORS} base kernels and 1..{MAX_TERMS} terms "
                         f"(got {len(factors)}, {len(terms)}); no CPU fallback")
    prog = [float(len(factors)), float(len(terms))]
    for kind, grad_mode, hyp in factors:
        hyp = [float(v) for v in hyp]
        prog += [float(kind), float(grad_mode), float(len(hyp))] + hyp
    for coef, expo in terms:
        assert len(expo) == len(factors)
        prog += [float(coef)] + [float(e) for e in expo]
    if ranges is not None:
        assert len(ranges) == len(factors)
        for start, count in ranges:
            prog += [float(start), float(count)]
    return make_spec(KERNEL_COMPOSITE, prog, noise, GRAD_REFERENCE)


class Context:
    """One GPU (dgp_ctx).  Not thread-safe, like the reference model."""

    def __init__(self, device: int | None = None):
        self.lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        st = self.lib.dgp_ctx_create(int(device), C.byref(h))
        if st != DGP_OK:
            raise DgpError(st, f"dgp_ctx_create(device={device}) failed: a B200 (sm_100) GPU is required; "
                               "there is no CPU fallback")
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dgp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------------------
    def check(self, st: int):
        if st == DGP_OK:
            return
        msg = self.lib.dgp_last_error(self.h).decode(errors="replace")
        if st == DGP_ERR_NOT_PD:
            raise NotPositiveDefinite(st, msg)
        raise DgpError(st, msg)

    def set_option(self, key: str, value: int):
        self.check(self.lib.dgp_ctx_set_option(self.h, key.encode(), int(value)))

    def trim(self):
        """Give the workspace arena back to the driver (handles stay valid)."""
        self.check(self.lib.dgp_ctx_trim(self.h))

    def info(self) -> dict:
        out = (C.c_int64 * 4)()
        self.check(self.lib.dgp_ctx_info(self.h, out, 4))
        return {"device": int(out[0]), "sm_count": int(out[1]), "mem_total": int(out[2]), "mem_free": int(out[3])}

    def timings(self) -> dict:
        out = np.zeros(16)
        self.check(self.lib.dgp_get_timings(self.h, _ptr(out), 16))
        names = ["assemble", "potrf", "solves", "inverse", "trace", "cross_assemble", "trsm", "syrk", "total",
                 "batch_total"]
        return {n: float(out[i]) for i, n in enumerate(names)}

    # -- data -------------------------------------------------------------------------------
    def data_create(self, X, y, mean=None):
        X = _f64(X)
        if X.ndim != 2:
            raise ValueError("X must be N x d")
        y = _f64(y).ravel()
        m = None if mean is None else _f64(mean).ravel()
        if y.size != X.shape[0] or (m is not None and m.size != y.size):
            raise ValueError("X, y, mean disagree on N")
        h = C.c_void_p()
        self.check(self.lib.dgp_data_create(self.h, _ptr(X), X.shape[0], X.shape[1], DGP_ROW_MAJOR, _ptr(y), _ptr(m),
                                            C.byref(h)))
        return h

    def data_destroy(self, h):
        if h:
            self.lib.dgp_data_destroy(self.h, h)

    # -- kernel matrices --------------------------------------------------------------------
    def kernel_matrix(self, spec, X1, X2=None, add_noise=False):
        s, keep = spec
        X1 = _f64(X1)
        n1, d = X1.shape
        if X2 is None:
            out = np.empty((n1, n1), order="F")
            st = self.lib.dgp_kernel_matrix(self.h, C.byref(s), _ptr(X1), n1, None, 0, d, DGP_ROW_MAJOR, _ptr(out), n1,
                                            int(bool(add_noise)))
        else:
            X2 = _f64(X2)
            if X2.shape[1] != d:
                raise ValueError("X1 and X2 disagree on d")
            n2 = X2.shape[0]
            out = np.empty((n1, n2), order="F")
            st = self.lib.dgp_kernel_matrix(self.h, C.byref(s), _ptr(X1), n1, _ptr(X2), n2, d, DGP_ROW_MAJOR, _ptr(out),
                                            n1, 0)
        self.check(st)
        return out

    def data_kernel_block(self, data, spec, row0: int, nrows: int, col0: int, ncols: int, add_noise=True):
        """K[row0:row0+nrows, col0:col0+ncols] of the training kernel matrix of a resident data set."""
        s, keep = spec
        out = np.empty((nrows, ncols), order="F")
        self.check(self.lib.dgp_data_kernel_block(self.h, data, C.byref(s), row0, nrows, col0, ncols, int(bool(add_noise)),
                                                  _ptr(out), nrows))
        return out

    def kernel_grad_matrix(self, spec, X1, X2=None, add_noise=False):
        """(K, [dK/dtheta ...]) in the order of energy_grad: the spec's hyper-parameters, then the noise level."""
        s, keep = spec
        X1 = _f64(X1)
        n1, d = X1.shape
        ng = int(self.lib.dgp_spec_grad_len(C.byref(s)))
        if ng < 0:
            raise DgpError(-ng, "malformed kernel spec")
        X2c = None if X2 is None else _f64(X2)
        n2 = n1 if X2c is None else X2c.shape[0]
        out = np.empty((1 + ng, n2, n1))          # plane-major, each plane column-major n1 x n2
        self.check(self.lib.dgp_kernel_grad_matrix(self.h, C.byref(s), _ptr(X1), n1, _ptr(X2c), n2 if X2c is not None else 0,
                                                   d, DGP_ROW_MAJOR, int(bool(add_noise)), _ptr(out)))
        planes = [np.asfortranarray(out[q].T) for q in range(1 + ng)]
        return planes[0], planes[1:]

    # -- energies ---------------------------------------------------------------------------
    def energy(self, data, spec) -> float:
        s, keep = spec
        e = C.c_double()
        st = self.lib.dgp_energy(self.h, data, C.byref(s), C.cast(C.byref(e), _dp))
        if st == DGP_ERR_NOT_PD:
            return float("inf")  # AbstractGPRegressionModel.scala:531-533
        self.check(st)
        return e.value

    def energy_terms(self, data, spec):
        """(r' K^-1 r, sum log L_ii); raises NotPositiveDefinite when the Cholesky fails."""
        s, keep = spec
        q, l = C.c_double(), C.c_double()
        self.check(self.lib.dgp_energy_terms(self.h, data, C.byref(s), C.cast(C.byref(q), _dp), C.cast(C.byref(l), _dp)))
        return q.value, l.value

    def model_terms(self, h):
        q, l = C.c_double(), C.c_double()
        self.check(self.lib.dgp_model_terms(self.h, h, C.cast(C.byref(q), _dp), C.cast(C.byref(l), _dp)))
        return q.value, l.value

    def data_set_basis(self, data, psi):
        """GPBasisFuncRegressionModel: psi is n x m (row i = lowB basisFunc(x_i)); None removes the term."""
        if psi is None:
            self.check(self.lib.dgp_data_set_basis(self.h, data, None, 0))
            return
        psi = _f64(psi)
        self.check(self.lib.dgp_data_set_basis(self.h, data, _ptr(psi), int(psi.shape[1])))

    def model_set_test_basis(self, model, psi_s):
        psi_s = _f64(psi_s)
        self.check(self.lib.dgp_model_set_test_basis(self.h, model, _ptr(psi_s), int(psi_s.shape[0]),
                                                     int(psi_s.shape[1])))

    def model_set_cross_spec(self, model, spec):
        """General noise models: the covariance-only spec used for k(X, X*) and k(X*, X*)."""
        s, keep = spec
        self.check(self.lib.dgp_model_set_cross_spec(self.h, model, C.byref(s)))

    def model_solve(self, model, B):
        """(K + noise)^-1 B for the columns of B (n or n x k) with the factor kept by fit()."""
        B = np.asarray(B, dtype=np.float64)
        one = B.ndim == 1
        Bf = np.asfortranarray(B.reshape(B.shape[0], -1))
        out = np.empty_like(Bf, order="F")
        self.check(self.lib.dgp_model_solve(self.h, model, _ptr(Bf), int(Bf.shape[1]), _ptr(out)))
        return out[:, 0] if one else out

    def model_cross_apply(self, model, Xs, V):
        """k(X, X*)' V for the columns of V (n or n x k); returns n_test (x k)."""
        Xs = _f64(Xs)
        V = np.asarray(V, dtype=np.float64)
        one = V.ndim == 1
        Vf = np.asfortranarray(V.reshape(V.shape[0], -1))
        out = np.empty((Xs.shape[0], Vf.shape[1]), order="F")
        self.check(self.lib.dgp_model_cross_apply(self.h, model, _ptr(Xs), Xs.shape[0], DGP_ROW_MAJOR, _ptr(Vf),
                                                  int(Vf.shape[1]), _ptr(out)))
        return out[:, 0] if one else out

    def energy_grad(self, data, spec):
        s, keep = spec
        e = C.c_double()
        ng = int(self.lib.dgp_spec_grad_len(C.byref(s)))
        if ng < 0:
            raise DgpError(-ng, "malformed kernel spec")
        g = np.empty(ng)
        st = self.lib.dgp_energy_grad(self.h, data, C.byref(s), C.cast(C.byref(e), _dp), _ptr(g))
        if st == DGP_ERR_NOT_PD:
            return float("inf"), np.full(ng, np.nan)  # :243-251
        self.check(st)
        return e.value, g

    def energy_batch(self, data, specs):
        n = len(specs)
        arr = (KernelSpec * n)(*[s for s, _ in specs])
        e = np.empty(n)
        status = np.zeros(n, dtype=np.int32)
        self.check(self.lib.dgp_energy_batch(self.h, data, arr, n, _ptr(e), status.ctypes.data_as(C.POINTER(C.c_int32))))
        return e, status

    @staticmethod
    def energy_batch_multi(ctxs, datas, specs):
        """dgp_energy_batch_multi: candidate i on ctxs[i % len(ctxs)], all GPUs concurrently, one process."""
        n, nc = len(specs), len(ctxs)
        arr = (KernelSpec * n)(*[s for s, _ in specs])
        hc = (C.c_void_p * nc)(*[c.h for c in ctxs])
        hd = (C.c_void_p * nc)(*[d for d in datas])
        e = np.empty(n)
        status = np.zeros(n, dtype=np.int32)
        ctxs[0].check(ctxs[0].lib.dgp_energy_batch_multi(hc, hd, nc, arr, n, _ptr(e),
                                                         status.ctypes.data_as(C.POINTER(C.c_int32))))
        return e, status

    # -- fit / predict ----------------------------------------------------------------------
    def fit(self, data, spec):
        s, keep = spec
        h = C.c_void_p()
        self.check(self.lib.dgp_fit(self.h, data, C.byref(s), C.byref(h)))
        return h

    def model_destroy(self, h):
        if h:
            self.lib.dgp_model_destroy(self.h, h)

    def model_energy(self, h) -> float:
        e = C.c_double()
        self.check(self.lib.dgp_model_energy(self.h, h, C.cast(C.byref(e), _dp)))
        return e.value

    def predict(self, model, Xs, mean_s=None, want_cov=True, want_bars=False, sigma=1.0):
        Xs = _f64(Xs)
        M = Xs.shape[0]
        ms = None if mean_s is None else _f64(mean_s).ravel()
        mean = np.empty(M)
        cov = np.empty((M, M), order="F") if want_cov else None
        lo = np.empty(M) if want_bars else None
        hi = np.empty(M) if want_bars else None
        st = self.lib.dgp_predict(self.h, model, _ptr(Xs), M, DGP_ROW_MAJOR, _ptr(ms), _ptr(mean), _ptr(cov), M,
                                  _ptr(lo), _ptr(hi), float(sigma))
        self.check(st)
        return mean, cov, lo, hi

    # -- measurement / building blocks ------------------------------------------------------
    def measure_peaks(self) -> dict:
        out = np.zeros(9)
        self.check(self.lib.dgp_measure_peaks(self.h, _ptr(out), 9))
        return {"dmma_tflops": float(out[0]), "dfma_tflops": float(out[1]), "hbm_write_gbs": float(out[2]),
                "hbm_copy_gbs": float(out[3]), "hbm_read_gbs": float(out[4]), "dmma_tflops_sustained": float(out[5]),
                "mixed_dmma_tflops": float(out[6]), "mixed_dfma_tflops": float(out[7]),
                "dmma_tflops_gemm_pattern": float(out[8])}

    def launch_count(self) -> int:
        n = C.c_uint64()
        self.check(self.lib.dgp_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def outer_block(self, n: int) -> int:
        return int(self.lib.dgp_outer_block(self.h, int(n)))

    def time_phase(self, data, spec, phase: int, reps: int = 3) -> float:
        s, keep = spec
        ms = C.c_double()
        self.check(self.lib.dgp_time_phase(self.h, data, C.byref(s), phase, reps, C.cast(C.byref(ms), _dp)))
        return ms.value

    def test_gemm(self, A, B, C_in=None, alpha=1.0, beta=0.0, transA=False, transB=False, lower_only=False):
        A = np.asfortranarray(A, dtype=np.float64)
        B = np.asfortranarray(B, dtype=np.float64)
        M, K = (A.shape[1], A.shape[0]) if transA else A.shape
        K2, N = (B.shape[1], B.shape[0]) if transB else B.shape
        if K != K2:
            raise ValueError("inner dimensions disagree")
        Cm = np.zeros((M, N), order="F") if C_in is None else np.array(C_in, dtype=np.float64, order="F")
        ms = C.c_double()
        self.check(self.lib.dgp_test_gemm(self.h, int(transA), int(transB), M, N, K, alpha, _ptr(A), A.shape[0],
                                          _ptr(B), B.shape[0], beta, _ptr(Cm), M, int(lower_only),
                                          C.cast(C.byref(ms), _dp)))
        return Cm, ms.value

    def test_potrf(self, A, want_linv=False, want_ainv=False):
        L = np.array(A, dtype=np.float64, order="F")
        n = L.shape[0]
        linv = np.empty((n, n), order="F") if want_linv else None
        ainv = np.empty((n, n), order="F") if want_ainv else None
        ld, ms = C.c_double(), C.c_double()
        self.check(self.lib.dgp_test_potrf(self.h, n, _ptr(L), n, _ptr(linv), _ptr(ainv), C.cast(C.byref(ld), _dp),
                                           C.cast(C.byref(ms), _dp)))
        return L, linv, ainv, ld.value, ms.value

    # -- distributed ------------------------------------------------------------------------
    def dist_unique_id(self) -> bytes:
        buf = (C.c_char * 128)()
        st = self.lib.dgp_dist_unique_id(C.cast(buf, C.c_void_p))
        if st != DGP_OK:
            raise DgpError(st, "ncclGetUniqueId failed")
        return bytes(buf)

    def dist_init(self, uid: bytes | None, rank: int, world: int, grid_rows: int, grid_cols: int):
        if uid is None:
            self.check(self.lib.dgp_dist_init(self.h, None, rank, world, grid_rows, grid_cols))
            return
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self.check(self.lib.dgp_dist_init(self.h, C.cast(buf, C.c_void_p), rank, world, grid_rows, grid_cols))

    def dist_info(self) -> dict:
        out = (C.c_int32 * 6)()
        self.check(self.lib.dgp_dist_info(self.h, out, 6))
        return {"nccl_ranks": int(out[0]), "nccl_rank": int(out[1]), "grid": [int(out[2]), int(out[3])],
                "emulated": bool(out[4]), "nccl_version": int(out[5])}

    def dist_finalize(self):
        self.check(self.lib.dgp_dist_finalize(self.h))

    def dist_energy(self, data, spec, tile: int = 512):
        s, keep = spec
        e, ms = C.c_double(), C.c_double()
        st = self.lib.dgp_dist_energy(self.h, data, C.byref(s), tile, C.cast(C.byref(e), _dp), C.cast(C.byref(ms), _dp))
        if st == DGP_ERR_NOT_PD:
            return float("inf"), ms.value
        self.check(st)
        return e.value, ms.value


def dist_plan(n: int, tile: int, grid_rows: int, grid_cols: int, rank: int) -> dict:
    """Host-only block-cyclic layout of one rank (works without a GPU)."""
    out = (C.c_int64 * 5)()
    st = load_library().dgp_dist_plan(n, tile, grid_rows, grid_cols, rank, out)
    if st != DGP_OK:
        raise DgpError(st, "dgp_dist_plan: bad arguments")
    return {"tiles": out[0], "local_rows": out[1], "local_cols": out[2], "owned_lower_tiles": out[3],
            "local_bytes": out[4]}


def device_count() -> int:
    """GPUs visible to this process (0 without a driver)."""
    try:
        cudart = C.CDLL("libcudart.so")
    except OSError:
        try:
            cudart = C.CDLL("libcudart.so.12")
        except OSError:
            return 0
    n = C.c_int(0)
    return int(n.value) if cudart.cudaGetDeviceCount(C.byref(n)) == 0 else 0


_default_ctx: Context | None = None


def default_context() -> Context:
    """The process-wide context on cuda:LOCAL_RANK (one process per GPU)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx
