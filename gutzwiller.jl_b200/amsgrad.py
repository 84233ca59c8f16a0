"""gradient_descent / amsgrad / adaptive_gradient_descent -- host mirror of src/amsgrad.jl.

The loop itself runs inside the library (``gwz_adaptive_gradient_descent``): for the two built-in
objectives (``KineticVQMC``, ``LocalEnergyEvaluator``) no Python code runs between iterations; any
other objective with ``val_and_grad`` / ``val_err_and_grad`` methods is driven through a C callback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import check, lib
from .localenergy import LocalEnergyEvaluator
from .vqmc import KineticVQMC

_REASONS = ("iterations", "params", "gradient", "value", "errorbars")


@dataclass
class GradientDescentResult:
    """``GradientDescentResult`` (amsgrad.jl:9-25); ``rows()`` is the Tables.jl view (amsgrad.jl:49-68)."""
    fun: object
    hyperparameters: dict
    initial_params: np.ndarray
    iterations: int
    converged: bool
    reason: str
    param: np.ndarray
    value: np.ndarray
    error: np.ndarray
    gradient: np.ndarray
    first_moment: np.ndarray
    second_moment: np.ndarray
    param_delta: np.ndarray = field(repr=False, default=None)

    def rows(self):
        for i in range(self.iterations):
            yield dict(**self.hyperparameters, iter=i + 1, param=self.param[i], value=self.value[i],
                       error=self.error[i], gradient=self.gradient[i], first_moment=self.first_moment[i],
                       second_moment=self.second_moment[i], param_delta=self.param_delta[i])

    def __repr__(self):
        return (f"GradientDescentResult\n  iterations: {self.iterations}\n  converged: {self.converged} "
                f"({self.reason})\n  last value: {self.value[-1]} ± {self.error[-1]}\n"
                f"  last params: {self.param[-1]}")


def val_and_grad(f, *args, **kw):
    """Generic ``val_and_grad`` (exported by the reference, Gutzwiller.jl:16): dispatch on the first argument."""
    from .hamiltonian import GutzwillerAnsatz
    from .vqmc import KineticVQMCResult, val_and_grad_result
    if isinstance(f, KineticVQMCResult):
        return val_and_grad_result(f)
    if isinstance(f, (KineticVQMC, LocalEnergyEvaluator, GutzwillerAnsatz)):
        return f.val_and_grad(*args, **kw)
    return f.val_and_grad(*args, **kw)


def val_err_and_grad(f, *args, **kw):
    """Generic ``val_err_and_grad``; objectives without error bars return err = 0 (abstract.jl:59-62)."""
    from .vqmc import KineticVQMCResult, val_err_and_grad_result
    if isinstance(f, KineticVQMCResult):
        return val_err_and_grad_result(f)
    if hasattr(f, "val_err_and_grad"):
        return f.val_err_and_grad(*args, **kw)
    val, grad = val_and_grad(f, *args, **kw)
    return val, 0.0, grad


def adaptive_gradient_descent(use_amsgrad: bool, f, params, *, maxiter=100, verbose=False, α=0.01, alpha=None,
                              first_moment_init=None, second_moment_init=None, fix_params=None, early_stop=True,
                              grad_tol=None, param_tol=None, val_tol=None, err_tol=0.0, fkwargs=None, β1=0.1, β2=0.01,
                              beta1=None, beta2=None) -> GradientDescentResult:
    """``adaptive_gradient_descent(φ, ψ, f, p_init; ...)`` (amsgrad.jl:105-226).  ``params`` may be a previous
    :class:`GradientDescentResult` to continue from its last parameters and moments (amsgrad.jl:109-116)."""
    alpha = α if alpha is None else alpha
    beta1 = β1 if beta1 is None else beta1
    beta2 = β2 if beta2 is None else beta2
    fkwargs = dict(fkwargs or {})
    if isinstance(params, GradientDescentResult):
        prev = params
        params = prev.param[-1]
        first_moment_init = prev.first_moment[-1] if first_moment_init is None else first_moment_init
        second_moment_init = prev.second_moment[-1] if second_moment_init is None else second_moment_init
    p0 = np.ascontiguousarray(np.atleast_1d(np.asarray(params, dtype=np.float64)).ravel())
    n = p0.size
    tol = math.sqrt(np.finfo(np.float64).eps)
    o = _capi.GdOpts()
    check(lib().gwz_gd_opts_default(C.byref(o)))
    o.maxiter, o.use_amsgrad, o.early_stop = int(maxiter), int(bool(use_amsgrad)), int(bool(early_stop))
    o.alpha, o.beta1, o.beta2 = float(alpha), float(beta1), float(beta2)
    o.grad_tol = tol if grad_tol is None else float(grad_tol)
    o.param_tol = tol if param_tol is None else float(param_tol)
    o.val_tol = tol if val_tol is None else float(val_tol)
    o.err_tol = float(err_tol)
    keep = []

    def _vec(x, ctype, dtype):
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=dtype)).ravel())
        if a.size != n:
            raise ValueError("moment / fix_params length does not match the number of parameters")
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(ctype))

    if first_moment_init is not None:
        o.first_moment_init = _vec(first_moment_init, C.c_double, np.float64)
    if second_moment_init is not None:
        o.second_moment_init = _vec(second_moment_init, C.c_double, np.float64)
    if fix_params is not None:
        o.fix_params = _vec(np.asarray(fix_params, dtype=bool), C.c_uint8, np.uint8)

    rows = int(maxiter) + 1
    bufs = {k: np.zeros((rows, n)) for k in ("param", "gradient", "first_moment", "second_moment", "param_delta")}
    bufs["value"], bufs["error"] = np.zeros(rows), np.zeros(rows)
    tr = _capi.GdTrace()
    for k, a in bufs.items():
        setattr(tr, k, a.ctypes.data_as(C.POINTER(C.c_double)))
    p0c = p0.ctypes.data_as(C.POINTER(C.c_double))

    # The in-library loops take the handles of the bit-level HubbardReal1D + GutzwillerAnsatz engine ONLY (a gwz_walkers
    # / gwz_sweep); every other engine's handle (generic ansaetze, stored operators, vector ansatz) goes through the
    # callback path below.  The library checks the handle's type tag as well.
    from .vqmc import Walkers
    if isinstance(f, KineticVQMC) and not fkwargs and f.result.record is not True and n == 1 \
            and type(f.result.walkers_handle) is Walkers:
        res = f.result
        pooled = res.estimator == "pooled"
        flags = _capi.RUN_POOLED if pooled else _capi.RUN_BLOCKING  # val_err_and_grad(qmc, p): state.jl:150-164
        last = _capi.VqmcResult()
        check(lib().gwz_amsgrad_vqmc(res.walkers_handle.handle, p0c, C.byref(o), f.steps, res.burn_in, res.kernel,
                                     flags, C.byref(tr), C.byref(last)))
        # what the reference leaves behind in qmc.result: the last evaluation
        res.steps, res.last = f.steps, last
        res.params = bufs["param"][max(tr.iterations - 1, 0)].copy()
        res.has_device_series = not pooled
        res.blocking_summary = last.blocking if not pooled else None
    elif isinstance(f, LocalEnergyEvaluator) and not fkwargs and n == 1 and f.is_bit_level():
        check(lib().gwz_amsgrad_sweep(f.handle, p0c, C.byref(o), C.byref(tr)))
    else:
        errors = []

        def _wrap(with_err):
            def cb(_user, p, np_, val, err, grad):
                try:
                    pv = np.array([p[i] for i in range(np_)])
                    if with_err:
                        v, e, g = val_err_and_grad(f, pv, **fkwargs)
                    else:
                        v, g = val_and_grad(f, pv, **fkwargs)
                        e = 0.0
                    val[0], err[0] = float(v), float(e)
                    g = np.atleast_1d(g)
                    for i in range(np_):
                        grad[i] = float(g[i])
                    return 0
                except Exception as exc:  # surfaced after the C loop returns
                    errors.append(exc)
                    return _capi.ERR_INVALID
            return _capi.OBJECTIVE_FN(cb)

        cb1, cb2 = _wrap(False), _wrap(True)
        rc = lib().gwz_adaptive_gradient_descent(cb1, cb2, None, p0c, n, C.byref(o), C.byref(tr))
        if errors:
            raise errors[0]
        check(rc)

    it = tr.iterations
    hyper = dict(α=alpha, β1=beta1, β2=beta2) if use_amsgrad else dict(α=alpha)
    return GradientDescentResult(f, hyper, p0, it, bool(tr.converged), _REASONS[tr.reason],
                                 **{k: v[:it].copy() for k, v in bufs.items()})


def gradient_descent(f, params, **kw) -> GradientDescentResult:
    """``gradient_descent(f, p0; ...)`` (amsgrad.jl:240-244): φ(m1,g)=g, ψ(m2,g)=1."""
    return adaptive_gradient_descent(False, f, params, **kw)


def amsgrad(f, params, **kw) -> GradientDescentResult:
    """``amsgrad(f, p0; β1=0.1, β2=0.01, ...)`` (amsgrad.jl:259-263)."""
    return adaptive_gradient_descent(True, f, params, **kw)
