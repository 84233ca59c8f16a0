"""ctypes binding of the CPU oracle (oracle/gutz_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "libgutz_oracle.so")
MAXW = 4


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("gutz_oracle.c", "gutz_oracle_ansatz.c", "gutz_oracle_sparse.c", "gutz_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


class Addr(C.Structure):
    _fields_ = [("w", C.c_uint64 * MAXW)]


class Model(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("u", C.c_double), ("t", C.c_double),
                ("B", C.c_int), ("W", C.c_int)]


class BlockingResult(C.Structure):
    _fields_ = [("mean", C.c_double), ("err", C.c_double), ("err_err", C.c_double),
                ("p_cov", C.c_double), ("k", C.c_int), ("blocks", C.c_int64)]


OBJECTIVE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                        C.POINTER(C.c_double), C.POINTER(C.c_double))


class GdOpts(C.Structure):
    _fields_ = [("maxiter", C.c_int), ("alpha", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("use_amsgrad", C.c_int), ("early_stop", C.c_int), ("grad_tol", C.c_double),
                ("param_tol", C.c_double), ("val_tol", C.c_double), ("err_tol", C.c_double),
                ("first_moment_init", C.POINTER(C.c_double)), ("second_moment_init", C.POINTER(C.c_double)),
                ("fix_params", C.POINTER(C.c_uint8))]


class Ansatz(C.Structure):
    _fields_ = [("m", Model), ("v", C.c_double), ("kind", C.c_int), ("P", C.c_int), ("normalization", C.c_double),
                ("kind1", C.c_int), ("kind2", C.c_int), ("norm1", C.c_int), ("norm2", C.c_int)]


ANSATZ_KINDS = {"gutzwiller": 0, "ext_gutzwiller": 1, "multinomial": 2, "density_profile": 3,
                "relative_jastrow": 4, "jastrow": 5}


class LeeObjective(C.Structure):
    _fields_ = [("m", C.POINTER(Model)), ("basis", C.c_void_p), ("dim", C.c_int64), ("mode", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        ap, mp = C.POINTER(Addr), C.POINTER(Model)
        L.go_model_init.argtypes = [mp, C.c_int, C.c_int, C.c_double, C.c_double]
        L.go_from_onr.argtypes = [mp, ip, ap]
        L.go_to_onr.argtypes = [mp, ap, ip]
        L.go_near_uniform.argtypes = [mp, ap]
        L.go_diagonal_element.argtypes = [mp, ap]
        L.go_diagonal_element.restype = C.c_double
        L.go_num_offdiagonals.argtypes = [mp, ap]
        L.go_get_offdiagonal.argtypes = [mp, ap, C.c_int, ap]
        L.go_get_offdiagonal.restype = C.c_double
        L.go_build_basis.argtypes = [mp, ap, C.c_void_p, C.c_int64]
        L.go_build_basis.restype = C.c_int64
        L.go_dimension.argtypes = [mp]
        L.go_dimension.restype = C.c_double
        L.go_ansatz_val.argtypes = [mp, ap, C.c_double]
        L.go_ansatz_val.restype = C.c_double
        L.go_ansatz_val_and_grad.argtypes = [mp, ap, C.c_double, dp, dp]
        L.go_kinetic_sample.argtypes = [mp, C.c_double, ap, C.c_double, dp, ap, dp, dp, dp, ip, ip]
        L.go_hop_rates.argtypes = [mp, C.c_double, ap, dp, dp]
        L.go_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.go_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.go_u01.restype = C.c_double
        L.go_kinetic_vqmc.argtypes = [mp, C.c_double, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64,
                                      C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.go_state_val_and_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, dp, dp]
        L.go_result_val_and_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, dp, dp]
        L.go_lee_val_and_grad.argtypes = [mp, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int, dp, dp]
        L.go_lee_val.argtypes = [mp, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_int]
        L.go_lee_val.restype = C.c_double
        L.go_lee_terms.argtypes = [mp, ap, C.c_double, dp]
        L.go_resample.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        L.go_blocking_analysis.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.POINTER(BlockingResult)]
        L.go_state_val_err_and_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, dp, dp, dp]
        L.go_adaptive_gradient_descent.argtypes = [OBJECTIVE, OBJECTIVE, C.c_void_p, dp, C.c_int,
                                                   C.POINTER(GdOpts), dp, dp, dp, dp, dp, dp, dp, ip, ip]
        L.go_lee_objective_val_and_grad.argtypes = [C.c_void_p, dp, C.c_int, dp, dp, dp]
        anp = C.POINTER(Ansatz)
        L.go_ansatz_init.argtypes = [anp, mp, C.c_double, C.c_int, C.c_int]
        L.go_ansatz_init_combination.argtypes = [anp, mp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
        L.go_gen_diagonal.argtypes = [anp, ap]
        L.go_gen_diagonal.restype = C.c_double
        L.go_gen_val_and_grad.argtypes = [anp, ap, C.c_void_p, dp, C.c_void_p]
        L.go_gen_kinetic_sample.argtypes = [anp, C.c_void_p, ap, C.c_double, C.c_void_p, ap, dp, dp, C.c_void_p, ip, ip]
        L.go_gen_hop_rates.argtypes = [anp, C.c_void_p, ap, C.c_void_p]
        L.go_gen_lee_terms.argtypes = [anp, ap, C.c_void_p, C.c_void_p]
        L.go_gen_lee_val_and_grad.argtypes = [anp, C.c_void_p, C.c_int64, C.c_void_p, dp, C.c_void_p]
        L.go_gen_kinetic_vqmc.argtypes = [anp, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(C.c_int64)]
        L.go_kinetic_vqmc_fast.argtypes = [mp, C.c_double, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        vp = C.c_void_p
        L.go_sparse_kinetic_sample.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_uint32, C.c_double, vp,
                                               C.POINTER(C.c_uint32), dp, dp, vp, ip, ip]
        L.go_sparse_kinetic_vqmc.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int64, C.c_int64, C.c_uint64,
                                             C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(C.c_int64),
                                             C.c_uint64]
        L.go_sparse_lee_terms.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_uint64, vp]
        L.go_sparse_lee_terms.restype = None
        L.go_sparse_lee_val_and_grad.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_uint64, dp, vp]
        L.go_sparse_lee_val_and_grad.restype = None
        _lib = L
    return _lib


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """HubbardReal1D(BoseFS{N,M}; u, t) + GutzwillerAnsatz, CPU restatement."""

    def __init__(self, N: int, M: int, u: float = 1.0, t: float = 1.0):
        self.L = lib()
        self.m = Model()
        rc = self.L.go_model_init(C.byref(self.m), N, M, u, t)
        if rc:
            raise ValueError(f"go_model_init failed rc={rc}")
        self.N, self.M, self.u, self.t = N, M, u, t
        self.B, self.W = self.m.B, self.m.W

    # -- addresses are numpy uint64[MAXW] rows (word 0 = LSB) --
    def _addr(self, words) -> Addr:
        a = Addr()
        w = np.asarray(words, dtype=np.uint64).ravel()
        for i in range(min(MAXW, len(w))):
            a.w[i] = int(w[i])
        return a

    @staticmethod
    def _words(a: Addr) -> np.ndarray:
        return np.array([a.w[i] for i in range(MAXW)], dtype=np.uint64)

    def from_onr(self, onr) -> np.ndarray:
        arr = (C.c_int * self.M)(*[int(x) for x in onr])
        a = Addr()
        self.L.go_from_onr(C.byref(self.m), arr, C.byref(a))
        return self._words(a)

    def to_onr(self, words) -> list:
        arr = (C.c_int * self.M)()
        a = self._addr(words)
        self.L.go_to_onr(C.byref(self.m), C.byref(a), arr)
        return list(arr)

    def near_uniform(self) -> np.ndarray:
        a = Addr()
        self.L.go_near_uniform(C.byref(self.m), C.byref(a))
        return self._words(a)

    def diagonal_element(self, words) -> float:
        a = self._addr(words)
        return self.L.go_diagonal_element(C.byref(self.m), C.byref(a))

    def num_offdiagonals(self, words) -> int:
        a = self._addr(words)
        return self.L.go_num_offdiagonals(C.byref(self.m), C.byref(a))

    def get_offdiagonal(self, words, chosen: int):
        a, out = self._addr(words), Addr()
        me = self.L.go_get_offdiagonal(C.byref(self.m), C.byref(a), chosen, C.byref(out))
        return self._words(out), me

    def offdiagonals(self, words):
        return [self.get_offdiagonal(words, c) for c in range(1, self.num_offdiagonals(words) + 1)]

    def dimension(self) -> int:
        return int(self.L.go_dimension(C.byref(self.m)))

    def build_basis(self, start=None) -> np.ndarray:
        start = self.near_uniform() if start is None else start
        cap = self.dimension()
        basis = np.zeros((cap, MAXW), dtype=np.uint64)
        a = self._addr(start)
        dim = self.L.go_build_basis(C.byref(self.m), C.byref(a), _vp(basis), cap)
        if dim < 0:
            raise RuntimeError("build_basis overflow")
        return basis[:dim]

    def ansatz(self, words, g: float) -> float:
        a = self._addr(words)
        return self.L.go_ansatz_val(C.byref(self.m), C.byref(a), g)

    def ansatz_val_and_grad(self, words, g: float):
        a = self._addr(words)
        v, d = C.c_double(), C.c_double()
        self.L.go_ansatz_val_and_grad(C.byref(self.m), C.byref(a), g, C.byref(v), C.byref(d))
        return v.value, d.value

    def kinetic_sample(self, words, g: float, u01: float):
        a, new = self._addr(words), Addr()
        buf = (C.c_double * (2 * self.M))()
        tau, el, gr = C.c_double(), C.c_double(), C.c_double()
        ch, nh = C.c_int(), C.c_int()
        rc = self.L.go_kinetic_sample(C.byref(self.m), g, C.byref(a), u01, buf, C.byref(new), C.byref(tau),
                                      C.byref(el), C.byref(gr), C.byref(ch), C.byref(nh))
        if rc:
            raise FloatingPointError("Infinite residence time")
        return dict(new_addr=self._words(new), tau=tau.value, eloc=el.value, grad=gr.value,
                    chosen=ch.value, n_hops=nh.value, cumsum=np.array(buf[:nh.value]))

    def hop_rates(self, words, g: float):
        a = self._addr(words)
        r = (C.c_double * (2 * self.M))()
        me = (C.c_double * (2 * self.M))()
        h = self.L.go_hop_rates(C.byref(self.m), g, C.byref(a), r, me)
        return np.array(r[:h]), np.array(me[:h])

    def u01(self, seed: int, walker: int, step: int) -> float:
        return self.L.go_u01(seed, walker, step)

    def kinetic_vqmc(self, g: float, addrs: np.ndarray, steps: int, seed: int = 0, walker0: int = 0,
                     step0: int = 0, n_threads: int = 0, series: bool = False):
        """addrs: uint64[n_walkers, MAXW], updated in place. Returns dict."""
        addrs = np.ascontiguousarray(addrs, dtype=np.uint64)
        nw = addrs.shape[0]
        val_w = np.zeros(nw)
        grad_w = np.zeros(nw)
        hops = C.c_int64()
        if series:
            sa = np.zeros((nw, steps, MAXW), dtype=np.uint64)
            st = np.zeros((nw, steps))
            se = np.zeros((nw, steps))
            sg = np.zeros((nw, steps))
            ptrs = [_vp(sa), _vp(st), _vp(se), _vp(sg)]
        else:
            sa = st = se = sg = None
            ptrs = [None, None, None, None]
        rc = self.L.go_kinetic_vqmc(C.byref(self.m), g, _vp(addrs), nw, steps, seed, walker0, step0, n_threads,
                                    *ptrs, _vp(val_w), _vp(grad_w), C.byref(hops))
        if rc:
            raise FloatingPointError("Infinite residence time")
        v, gr = C.c_double(), C.c_double()
        self.L.go_result_val_and_grad(_vp(val_w), _vp(grad_w), nw, C.byref(v), C.byref(gr))
        return dict(addrs=addrs, val_w=val_w, grad_w=grad_w, val=v.value, grad=gr.value, hops=hops.value,
                    series_addr=sa, series_tau=st, series_eloc=se, series_grad=sg)

    def kinetic_vqmc_fast(self, g: float, addrs: np.ndarray, steps: int, seed: int = 0, walker0: int = 0,
                          step0: int = 0, n_threads: int = 0):
        """Tuned CPU code for the same walk (closed-form ratios): an upper bound of the CPU side, not the reference."""
        addrs = np.ascontiguousarray(addrs, dtype=np.uint64)
        nw = addrs.shape[0]
        val_w, grad_w, hops = np.zeros(nw), np.zeros(nw), C.c_int64()
        rc = self.L.go_kinetic_vqmc_fast(C.byref(self.m), g, _vp(addrs), nw, steps, seed, walker0, step0, n_threads,
                                         _vp(val_w), _vp(grad_w), C.byref(hops))
        if rc:
            raise FloatingPointError("Infinite residence time")
        return dict(addrs=addrs, val_w=val_w, grad_w=grad_w, val=val_w.mean(), grad=grad_w.mean(), hops=hops.value)

    def lee_val_and_grad(self, basis: np.ndarray, g: float, mode: int = 0, n_threads: int = 0):
        basis = np.ascontiguousarray(basis, dtype=np.uint64)
        v, gr = C.c_double(), C.c_double()
        self.L.go_lee_val_and_grad(C.byref(self.m), _vp(basis), basis.shape[0], g, mode, n_threads,
                                   C.byref(v), C.byref(gr))
        return v.value, gr.value

    def lee_val(self, basis: np.ndarray, g: float, mode: int = 0, n_threads: int = 0) -> float:
        basis = np.ascontiguousarray(basis, dtype=np.uint64)
        return self.L.go_lee_val(C.byref(self.m), _vp(basis), basis.shape[0], g, mode, n_threads)

    def lee_terms(self, words, g: float) -> np.ndarray:
        a = self._addr(words)
        out = (C.c_double * 4)()
        self.L.go_lee_terms(C.byref(self.m), C.byref(a), g, out)
        return np.array(out[:])

    def amsgrad_lee(self, basis: np.ndarray, p0: float, maxiter=100, alpha=0.01, beta1=0.1, beta2=0.01,
                    use_amsgrad=True, early_stop=True, mode=0, m1_init=None, m2_init=None):
        basis = np.ascontiguousarray(basis, dtype=np.uint64)
        obj = LeeObjective(C.pointer(self.m), _vp(basis), basis.shape[0], mode)
        eps = float(np.sqrt(np.finfo(np.float64).eps))
        o = GdOpts(maxiter, alpha, beta1, beta2, int(use_amsgrad), int(early_stop), eps, eps, eps, 0.0,
                   None, None, None)
        keep = []
        if m1_init is not None:
            a = (C.c_double * 1)(m1_init); keep.append(a); o.first_moment_init = a
        if m2_init is not None:
            a = (C.c_double * 1)(m2_init); keep.append(a); o.second_moment_init = a
        rows = maxiter + 1
        tr = {k: np.zeros(rows) for k in ("param", "value", "error", "gradient", "first_moment",
                                            "second_moment", "param_delta")}
        conv, reason = C.c_int(), C.c_int()
        dp = C.POINTER(C.c_double)
        fn = C.cast(self.L.go_lee_objective_val_and_grad, OBJECTIVE)
        p0a = (C.c_double * 1)(p0)
        n = self.L.go_adaptive_gradient_descent(
            fn, fn, C.cast(C.pointer(obj), C.c_void_p), p0a, 1, C.byref(o),
            tr["param"].ctypes.data_as(dp), tr["value"].ctypes.data_as(dp), tr["error"].ctypes.data_as(dp),
            tr["gradient"].ctypes.data_as(dp), tr["first_moment"].ctypes.data_as(dp),
            tr["second_moment"].ctypes.data_as(dp), tr["param_delta"].ctypes.data_as(dp),
            C.byref(conv), C.byref(reason))
        out = {k: v[:n].copy() for k, v in tr.items()}
        out.update(iterations=n, converged=bool(conv.value),
                   reason=["iterations", "params", "gradient", "value", "errorbars"][reason.value])
        return out

    def resample(self, times, values, length=None) -> np.ndarray:
        times = np.ascontiguousarray(times, dtype=np.float64)
        values = np.ascontiguousarray(values, dtype=np.float64)
        if len(times) != len(values):
            raise ValueError("Lengths of `residence_times` and `values` do not match!")
        length = len(values) if length is None else length
        out = np.zeros(length)
        self.L.go_resample(_vp(out), length, _vp(times), _vp(values), len(values))
        return out

    def blocking_analysis(self, v) -> BlockingResult:
        v = np.ascontiguousarray(v, dtype=np.float64)
        r = BlockingResult()
        self.L.go_blocking_analysis(_vp(v), len(v), 0.01, C.byref(r))
        return r


class AnsatzOracle:
    """(Extended)HubbardReal1D + one of the reference's single-component ansaetze, CPU restatement
    (oracle/gutz_oracle_ansatz.c).  `base` is an Oracle (model + address helpers)."""

    def __init__(self, base: Oracle, kind, v: float = 0.0, normalize=False):
        """kind: a name from ANSATZ_KINDS, or a pair of names for CombinationAnsatz (a1 + a2, combination.jl);
        normalize: bool, or a pair of bools for the two components."""
        self.base, self.L = base, base.L
        self.a = Ansatz()
        if isinstance(kind, (tuple, list)):
            n1, n2 = normalize if isinstance(normalize, (tuple, list)) else (normalize, normalize)
            rc = self.L.go_ansatz_init_combination(C.byref(self.a), C.byref(base.m), v, ANSATZ_KINDS[kind[0]], int(n1),
                                                   ANSATZ_KINDS[kind[1]], int(n2))
            kind = "+".join(kind)
        else:
            rc = self.L.go_ansatz_init(C.byref(self.a), C.byref(base.m), v, ANSATZ_KINDS[kind], int(normalize))
        if rc:
            raise ValueError(f"unknown ansatz kind {kind}")
        self.kind, self.v, self.P, self.M = kind, v, self.a.P, base.M

    def _p(self, params) -> np.ndarray:
        p = np.ascontiguousarray(np.atleast_1d(params), dtype=np.float64)
        if p.shape != (self.P,):
            raise ValueError(f"expected {self.P} parameters")
        return p

    def diagonal_element(self, words) -> float:
        a = self.base._addr(words)
        return self.L.go_gen_diagonal(C.byref(self.a), C.byref(a))

    def val(self, words, params) -> float:
        a, p, v = self.base._addr(words), self._p(params), C.c_double()
        self.L.go_gen_val_and_grad(C.byref(self.a), C.byref(a), _vp(p), C.byref(v), None)
        return v.value

    def val_and_grad(self, words, params):
        a, p, v, g = self.base._addr(words), self._p(params), C.c_double(), np.zeros(self.P)
        self.L.go_gen_val_and_grad(C.byref(self.a), C.byref(a), _vp(p), C.byref(v), _vp(g))
        return v.value, g

    def kinetic_sample(self, words, params, u01: float):
        a, new, p = self.base._addr(words), Addr(), self._p(params)
        buf, gr = np.zeros(2 * self.M), np.zeros(self.P)
        tau, el, ch, nh = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        rc = self.L.go_gen_kinetic_sample(C.byref(self.a), _vp(p), C.byref(a), u01, _vp(buf), C.byref(new),
                                          C.byref(tau), C.byref(el), _vp(gr), C.byref(ch), C.byref(nh))
        if rc:
            raise FloatingPointError("Infinite residence time")
        return dict(new_addr=self.base._words(new), tau=tau.value, eloc=el.value, grad=gr, chosen=ch.value,
                    n_hops=nh.value, cumsum=buf[:nh.value].copy())

    def hop_rates(self, words, params) -> np.ndarray:
        a, p, r = self.base._addr(words), self._p(params), np.zeros(2 * self.M)
        h = self.L.go_gen_hop_rates(C.byref(self.a), _vp(p), C.byref(a), _vp(r))
        return r[:h].copy()

    def lee_terms(self, words, params) -> np.ndarray:
        a, p, out = self.base._addr(words), self._p(params), np.zeros(2 + 2 * self.P)
        self.L.go_gen_lee_terms(C.byref(self.a), C.byref(a), _vp(p), _vp(out))
        return out

    def lee_val_and_grad(self, basis: np.ndarray, params):
        basis, p = np.ascontiguousarray(basis, dtype=np.uint64), self._p(params)
        v, g = C.c_double(), np.zeros(self.P)
        self.L.go_gen_lee_val_and_grad(C.byref(self.a), _vp(basis), basis.shape[0], _vp(p), C.byref(v), _vp(g))
        return v.value, g

    def kinetic_vqmc(self, params, addrs: np.ndarray, steps: int, seed: int = 0, walker0: int = 0, step0: int = 0,
                     n_threads: int = 0, series: bool = False):
        addrs, p = np.ascontiguousarray(addrs, dtype=np.uint64), self._p(params)
        nw = addrs.shape[0]
        val_w, grad_w, hops = np.zeros(nw), np.zeros((nw, self.P)), C.c_int64()
        st = np.zeros((nw, steps)) if series else None
        se = np.zeros((nw, steps)) if series else None
        sa = np.zeros((nw, steps, MAXW), dtype=np.uint64) if series else None
        rc = self.L.go_gen_kinetic_vqmc(C.byref(self.a), _vp(p), _vp(addrs), nw, steps, seed, walker0, step0,
                                        n_threads, _vp(st) if series else None, _vp(se) if series else None,
                                        _vp(sa) if series else None, _vp(val_w), _vp(grad_w), C.byref(hops))
        if rc:
            raise FloatingPointError("Infinite residence time")
        return dict(addrs=addrs, val_w=val_w, grad_w=grad_w, val=val_w.mean(), grad=grad_w.mean(axis=0),
                    hops=hops.value, series_tau=st, series_eloc=se, series_addr=sa)


class SparseOracle:
    """The hot path over a materialised operator (oracle/gutz_oracle_sparse.c): rows = offdiagonals in reference
    order as (col, melem), diag, and an ansatz table psi[dim] with grad_ratio[dim, P] = grad psi / psi."""

    def __init__(self, row_ptr, col, melem, diag):
        self.L = lib()
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.melem = np.ascontiguousarray(melem, dtype=np.float64)
        self.diag = np.ascontiguousarray(diag, dtype=np.float64)
        self.dim = self.diag.shape[0]
        assert self.row_ptr.shape == (self.dim + 1,) and self.col.shape == self.melem.shape == (int(self.row_ptr[-1]),)
        self.maxrow = int(np.diff(self.row_ptr.astype(np.int64)).max())
        self.set_table(np.ones(self.dim), None)

    def set_table(self, psi, grad_ratio=None):
        self.psi = np.ascontiguousarray(psi, dtype=np.float64)
        assert self.psi.shape == (self.dim,)
        if grad_ratio is None:
            self.P, self.gr = 0, np.zeros((self.dim, 0))
        else:
            self.gr = np.ascontiguousarray(np.asarray(grad_ratio, dtype=np.float64).reshape(self.dim, -1))
            self.P = self.gr.shape[1]

    def set_gutzwiller(self, g: float):
        """GutzwillerAnsatz (gutzwiller.jl:25-36): psi = exp(-g D), grad psi / psi = -D"""
        self.set_table(np.exp(-g * self.diag), -self.diag.reshape(-1, 1))

    def _ops(self):
        return (_vp(self.row_ptr), _vp(self.col), _vp(self.melem), _vp(self.diag), _vp(self.psi), _vp(self.gr), self.P)

    def kinetic_sample(self, k: int, u01: float):
        buf, gr = np.zeros(self.maxrow), np.zeros(max(self.P, 1))
        new, tau, el, ch, nh = C.c_uint32(), C.c_double(), C.c_double(), C.c_int(), C.c_int()
        rc = self.L.go_sparse_kinetic_sample(*self._ops(), int(k), u01, _vp(buf), C.byref(new), C.byref(tau),
                                             C.byref(el), _vp(gr), C.byref(ch), C.byref(nh))
        if rc:
            raise FloatingPointError("Infinite residence time")
        return dict(new_index=new.value, tau=tau.value, eloc=el.value, grad=gr[:self.P].copy(), chosen=ch.value,
                    n_hops=nh.value, cumsum=buf[:nh.value].copy())

    def kinetic_vqmc(self, idx, steps: int, seed: int = 0, walker0: int = 0, step0: int = 0, series: bool = False):
        idx = np.ascontiguousarray(idx, dtype=np.uint32).copy()
        nw = idx.shape[0]
        val_w, grad_w, hops = np.zeros(nw), np.zeros((nw, self.P)), C.c_int64()
        si = np.zeros((nw, steps), dtype=np.uint32) if series else None
        st = np.zeros((nw, steps)) if series else None
        se = np.zeros((nw, steps)) if series else None
        rc = self.L.go_sparse_kinetic_vqmc(*self._ops(), _vp(idx), nw, steps, seed, walker0, step0,
                                           _vp(si) if series else None, _vp(st) if series else None,
                                           _vp(se) if series else None, _vp(val_w), _vp(grad_w), C.byref(hops), self.dim)
        if rc:
            raise FloatingPointError("Infinite residence time")
        return dict(idx=idx, val_w=val_w, grad_w=grad_w, val=val_w.mean(), grad=grad_w.mean(axis=0), hops=hops.value,
                    series_idx=si, series_tau=st, series_eloc=se)

    def lee_terms(self, k: int) -> np.ndarray:
        out = np.zeros(2 + 2 * self.P)
        self.L.go_sparse_lee_terms(*self._ops(), int(k), _vp(out))
        return out

    def lee_val_and_grad(self):
        v, g = C.c_double(), np.zeros(max(self.P, 1))
        self.L.go_sparse_lee_val_and_grad(*self._ops(), self.dim, C.byref(v), _vp(g))
        return v.value, g[:self.P].copy()
