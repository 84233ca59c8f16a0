"""kinetic_vqmc / KineticVQMC / KineticVQMCResult -- host mirror of src/kinetic-vqmc/*.jl.

Same names, argument meaning and error behaviour as the reference; the walkers live in HBM behind
a ``gwz_walkers`` handle and every step runs in the persistent CUDA walker kernel.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import check, lib
from .hamiltonian import AbstractAnsatz, BoseFS, GutzwillerAnsatz, HubbardReal1D, as_params
from .runtime import Context, default_context, distributed_state, shard_range

DEFAULT_SEED = 0x9E3779B97F4A7C15
RECORD_LIMIT_BYTES = 256 << 20  # kinetic_vqmc ships per-step series + addresses to the host up to this size
DEVICE_SERIES_LIMIT_BYTES = 48 << 30  # beyond the host limit the (tau, E_loc) series stay in HBM up to this size


def default_walkers(samples: float) -> int:
    """GPU default for the `walkers` keyword.

    The reference derives its default from the thread count (qmc.jl:132).  Here it is the largest
    power of two that still leaves every walker >= 1024 steps (all walkers start at the same
    address and the reference has no burn-in, so short walkers bias the estimate), capped at 2^20.
    An explicit ``walkers=`` is always honoured exactly.
    """
    w = max(1, int(samples) // 1024)
    return min(1 << 20, 1 << (w.bit_length() - 1))


@dataclass
class BlockingResult:
    mean: float
    err: float
    err_err: float
    p_cov: float
    k: int
    blocks: int


@dataclass
class CombinedBlockingResult:
    """state.jl:238-256"""
    mean: float
    err: float
    err_err: float
    results: list


def resample(residence_times, values, len=None) -> np.ndarray:  # noqa: A002 - keyword name of the reference
    """``resample(residence_times, values; len)`` (state.jl:183-231)."""
    t = np.ascontiguousarray(residence_times, dtype=np.float64).ravel()
    v = np.ascontiguousarray(values, dtype=np.float64).ravel()
    if t.size != v.size:
        raise ValueError("Lengths of `residence_times` and `values` do not match!")  # state.jl:195-197
    n_out = t.size if len is None else int(len)
    out = np.zeros(n_out)
    check(lib().gwz_resample(out.ctypes.data_as(C.c_void_p), n_out, t.ctypes.data_as(C.c_void_p),
                             v.ctypes.data_as(C.c_void_p), t.size))
    return out


def blocking_analysis(v) -> BlockingResult:
    """Rimu.StatsTools.blocking_analysis (used at state.jl:53,273)."""
    v = np.ascontiguousarray(v, dtype=np.float64).ravel()
    r = _capi.BlockingResult()
    check(lib().gwz_blocking_analysis(v.ctypes.data_as(C.c_void_p), v.size, C.byref(r)))
    return BlockingResult(r.mean, r.err, r.err_err, r.p_cov, r.k, r.blocks)


class Walkers:
    """gwz_walkers: n walkers (global ids offset..offset+n-1) resident on one device."""

    def __init__(self, model, n_walkers: int, start_words, seed: int = DEFAULT_SEED, global_offset: int = 0):
        self.model = model
        self.n = int(n_walkers)
        self.offset = int(global_offset)
        self._h = C.c_void_p()
        sw = np.ascontiguousarray(start_words, dtype=np.uint64)
        check(lib().gwz_walkers_create(model.handle, self.n, self.offset, sw.ctypes.data_as(C.POINTER(C.c_uint64)),
                                       int(seed) & 0xFFFFFFFFFFFFFFFF, C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_walkers_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def steps_done(self) -> int:
        n, s = C.c_uint64(), C.c_uint64()
        check(lib().gwz_walkers_count(self._h, C.byref(n), C.byref(s)))
        return s.value

    def addresses(self) -> np.ndarray:
        out = np.zeros((self.n, self.model.words), dtype=np.uint64)
        check(lib().gwz_walkers_get_addresses(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_addresses(self, addrs) -> None:
        a = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(self.n, self.model.words)
        check(lib().gwz_walkers_set_addresses(self._h, a.ctypes.data_as(C.c_void_p)))

    def set_step_counter(self, steps_done: int) -> None:
        check(lib().gwz_walkers_set_step_counter(self._h, int(steps_done)))

    def run(self, params, steps: int, burn_in: int = 0, kernel: int = _capi.KERNEL_BITPARALLEL,
            keep_acc: bool = False, pooled: bool = False, series: bool = False,
            blocking: bool = False) -> _capi.VqmcResult:
        """gwz_vqmc_run.  series: keep the (tau, E_loc) series in HBM (GWZ_RUN_SERIES); blocking: also run the
        per-walker resample + blocking analysis on the device in the same call (r.blocking, GWZ_RUN_BLOCKING)."""
        p = as_params(params, 1)
        r = _capi.VqmcResult()
        flags = (_capi.RUN_KEEP_ACC if keep_acc else 0) | (_capi.RUN_POOLED if pooled else 0) | \
            (_capi.RUN_SERIES if series else 0) | (_capi.RUN_BLOCKING if blocking else 0)
        check(lib().gwz_vqmc_run(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), int(steps), int(burn_in),
                                 int(kernel), flags, C.byref(r)))
        return r

    def run_record(self, params, steps: int, kernel: int = _capi.KERNEL_BITPARALLEL, keep_acc: bool = False,
                   want_addresses: bool = True):
        p = as_params(params, 1)
        steps = int(steps)
        tau = np.zeros((steps, self.n))
        eloc = np.zeros((steps, self.n))
        addr = np.zeros((steps, self.n, self.model.words), dtype=np.uint64) if want_addresses else None
        r = _capi.VqmcResult()
        check(lib().gwz_vqmc_run_record(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), steps, int(kernel),
                                        _capi.RUN_KEEP_ACC if keep_acc else 0, tau.ctypes.data_as(C.c_void_p),
                                        eloc.ctypes.data_as(C.c_void_p),
                                        None if addr is None else addr.ctypes.data_as(C.c_void_p), C.byref(r)))
        return r, tau, eloc, addr

    def run_trace(self, params, steps: int, u01=None, keep_acc: bool = False):
        """gwz_vqmc_run_trace: recorded steps of the production kernel.  Returns (result, tau, eloc, hop), each
        [steps, n]; hop = 2 z + right (see the header).  ``u01`` ([steps, n]) replaces the Philox stream."""
        p = as_params(params, 1)
        steps = int(steps)
        tau, eloc = np.zeros((steps, self.n)), np.zeros((steps, self.n))
        hop = np.zeros((steps, self.n), dtype=np.int32)
        up = None
        if u01 is not None:
            u = np.ascontiguousarray(u01, dtype=np.float64).reshape(steps, self.n)
            up = u.ctypes.data_as(C.c_void_p)
        r = _capi.VqmcResult()
        check(lib().gwz_vqmc_run_trace(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), steps,
                                       _capi.RUN_KEEP_ACC if keep_acc else 0, up, tau.ctypes.data_as(C.c_void_p),
                                       eloc.ctypes.data_as(C.c_void_p), hop.ctypes.data_as(C.c_void_p), C.byref(r)))
        return r, tau, eloc, hop

    def sample_vector(self, params, steps: int, kernel: int = _capi.KERNEL_BITPARALLEL, keep_acc: bool = False,
                      cap: int | None = None):
        """``DVec(res)`` (state.jl:303-319) histogrammed on the device: (addresses uint64[k, 1], values float64[k])."""
        p = as_params(params, 1)
        cap = int(min(int(steps) * self.n, self.model.dimension())) if cap is None else int(cap)
        addr, val = np.zeros(cap, dtype=np.uint64), np.zeros(cap)
        cnt, r = C.c_uint64(), _capi.VqmcResult()
        check(lib().gwz_vqmc_sample_vector(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), int(steps), int(kernel),
                                           _capi.RUN_KEEP_ACC if keep_acc else 0, cap, addr.ctypes.data_as(C.c_void_p),
                                           val.ctypes.data_as(C.c_void_p), C.byref(cnt), C.byref(r)))
        return addr[: cnt.value].reshape(-1, 1), val[: cnt.value], r

    def estimates(self):
        v, g = np.zeros(self.n), np.zeros((self.n, 1))
        check(lib().gwz_walkers_estimates(self._h, v.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p)))
        return v, g

    def blocking(self, resample_len: int | None = None, per_walker: bool = False):
        """gwz_walkers_blocking: per-walker ``blocking_analysis(resample(tau, E_loc; len))`` of the series kept in
        HBM, combined over all walkers (state.jl:150-164, 265-285).  Returns the summary, and with ``per_walker``
        also a dict of arrays (mean, err, err_err, k, blocks)."""
        summ = _capi.BlockingSummary()
        arrs = None
        ptrs = [None] * 5
        if per_walker:
            arrs = dict(mean=np.zeros(self.n), err=np.zeros(self.n), err_err=np.zeros(self.n),
                        k=np.zeros(self.n, dtype=np.int32), blocks=np.zeros(self.n, dtype=np.int64))
            ptrs = [arrs[k].ctypes.data_as(C.c_void_p) for k in ("mean", "err", "err_err", "k", "blocks")]
        check(lib().gwz_walkers_blocking(self._h, int(resample_len or 0), C.byref(summ), *ptrs))
        return (summ, arrs) if per_walker else summ

    def series(self):
        """(tau, eloc), each [steps, n]: host copies of the series kept in HBM by ``run(..., series=True)``."""
        steps = C.c_uint64()
        check(lib().gwz_walkers_series(self._h, C.byref(steps), None, None))
        tau, eloc = np.zeros((steps.value, self.n)), np.zeros((steps.value, self.n))
        check(lib().gwz_walkers_series(self._h, C.byref(steps), tau.ctypes.data_as(C.c_void_p),
                                       eloc.ctypes.data_as(C.c_void_p)))
        return tau, eloc


def series_blocking(tau, eloc, resample_len: int | None = None, ctx: Context | None = None, resampled: bool = False):
    """gwz_series_blocking: the device kernel of ``val_err_and_grad`` on caller-supplied series.
    tau, eloc: [steps, n] (n independent series in the columns).  Returns (summary, dict of per-series arrays);
    with ``resampled=True`` the dict also holds the equal-time series the kernel analysed, [len, n]."""
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    eloc = np.ascontiguousarray(eloc, dtype=np.float64)
    if tau.ndim == 1:
        tau, eloc = tau[:, None].copy(), eloc[:, None].copy()
    if tau.shape != eloc.shape:
        raise ValueError("Lengths of `residence_times` and `values` do not match!")  # state.jl:195-197
    steps, n = tau.shape
    ctx = ctx or default_context()
    summ = _capi.BlockingSummary()
    arrs = dict(mean=np.zeros(n), err=np.zeros(n), err_err=np.zeros(n), k=np.zeros(n, dtype=np.int32),
                blocks=np.zeros(n, dtype=np.int64))
    rs = np.zeros((int(resample_len or steps), n)) if resampled else None
    check(lib().gwz_series_blocking(ctx.handle, tau.ctypes.data_as(C.c_void_p), eloc.ctypes.data_as(C.c_void_p), steps, n,
                                    int(resample_len or 0), C.byref(summ),
                                    *[arrs[k].ctypes.data_as(C.c_void_p) for k in ("mean", "err", "err_err", "k", "blocks")],
                                    None if rs is None else rs.ctypes.data_as(C.c_void_p)))
    if rs is not None:
        arrs["resampled"] = rs
    return summ, arrs


def run_group(walker_sets, params, steps: int, burn_in: int = 0, kernel: int = _capi.KERNEL_BITPARALLEL,
              keep_acc: bool = False) -> _capi.VqmcResult:
    """gwz_vqmc_run_group: several walker sets (logical shards / several devices of one process)."""
    p = as_params(params, 1)
    arr = (C.c_void_p * len(walker_sets))(*[w.handle for w in walker_sets])
    r = _capi.VqmcResult()
    check(lib().gwz_vqmc_run_group(arr, len(walker_sets), p.ctypes.data_as(C.POINTER(C.c_double)), int(steps),
                                   int(burn_in), int(kernel), _capi.RUN_KEEP_ACC if keep_acc else 0, C.byref(r)))
    return r


class KineticVQMCResult:
    """``KineticVQMCResult`` (state.jl:75-77): the walkers plus what has been sampled so far.

    ``val_and_grad(res)``, ``val_err_and_grad(res)``, ``mean(res)``, ``var(res)``,
    ``local_energy_estimator(res)`` and ``rows()`` follow state.jl:137-164,265-298,91-132.
    """

    def __init__(self, hamiltonian, ansatz, params, walkers: Walkers, record: bool, kernel: int, burn_in: int = 0,
                 estimator: str = "walker_mean"):
        if estimator not in ("walker_mean", "pooled"):
            raise ValueError("estimator must be 'walker_mean' (reference, state.jl:137-149) or 'pooled'")
        self.estimator = estimator
        self.hamiltonian, self.ansatz = hamiltonian, ansatz
        self.params = as_params(params, ansatz.N_PARAMS)
        self.walkers_handle = walkers
        self.record = record
        self.kernel = kernel
        self.burn_in = int(burn_in)
        self.last: _capi.VqmcResult | None = None
        self.blocking_summary = None      # gwz_blocking_summary of the last val_err_and_grad run, if any
        self.has_device_series = False    # the (tau, E_loc) series of the accumulated steps are in HBM
        self.steps = 0  # per walker
        self._tau: list[np.ndarray] = []
        self._eloc: list[np.ndarray] = []
        self._addr: list[np.ndarray] = []

    # -- sampling --
    def _device_series(self) -> bool:
        return self.record == "device" and isinstance(self.walkers_handle, Walkers)

    def _run(self, steps: int, keep: bool, want_err: bool | None = None):
        """want_err: the caller is ``val_err_and_grad``: keep the series in HBM and block them in the same call
        (None: whatever the record mode of this result says)."""
        steps = int(steps)
        w = self.walkers_handle
        burn = self.burn_in if self.steps == 0 or not keep else 0
        self.blocking_summary = None
        if self.record is not True and isinstance(w, Walkers) and self.estimator != "pooled" and \
                (want_err or (want_err is None and self.record == "device")):
            if keep and not self.has_device_series:
                raise RuntimeError("cannot extend the series: the earlier steps of this result were not recorded")
            r = w.run(self.params, steps, burn_in=burn, kernel=self.kernel, keep_acc=keep, series=True,
                      blocking=bool(want_err))
            self.has_device_series = True
            if want_err:
                self.blocking_summary = r.blocking
            self.steps = self.steps + steps if keep else steps
            self.last = r
            return self
        self.has_device_series = False
        if self.record is True:
            if burn:
                w.run(self.params, burn, kernel=self.kernel)  # thermalise; estimators are cleared next
            r, tau, eloc, addr = w.run_record(self.params, steps, kernel=self.kernel, keep_acc=keep)
            if not keep:
                self._tau, self._eloc, self._addr = [], [], []
            self._tau.append(tau)
            self._eloc.append(eloc)
            self._addr.append(addr)
        else:
            if burn:
                w.run(self.params, max(burn, 1), kernel=self.kernel)
            r = w.run(self.params, steps, kernel=self.kernel, keep_acc=keep, pooled=self.estimator == "pooled")
        self.steps = self.steps + steps if keep else steps
        self.last = r
        return self

    # -- series access (reference: per-walker vectors, state.jl:15-18) --
    @property
    def n_walkers(self) -> int:
        return int(self.last.walkers) if self.last is not None else self.walkers_handle.n

    def __len__(self):  # sum(length, res.states), the row count of the Tables.jl view
        return self.steps * self.walkers_handle.n

    def _need_series(self):
        if self.record is not True or not self._tau:
            raise RuntimeError("per-step series were not shipped to the host for this run (record=True does that)")

    def residence_times(self) -> np.ndarray:
        """[walker, step]"""
        self._need_series()
        return np.concatenate(self._tau, axis=0).T

    def local_energies(self) -> np.ndarray:
        self._need_series()
        return np.concatenate(self._eloc, axis=0).T

    def addresses(self) -> np.ndarray:
        """[walker, step, word]"""
        self._need_series()
        return np.concatenate(self._addr, axis=0).transpose(1, 0, 2)

    def rows(self):
        """Tables.jl rows (state.jl:91-132): (walker, step, address, residence_time, local_energy)."""
        tau, el, ad = self.residence_times(), self.local_energies(), self.addresses()
        N, M = self.hamiltonian.address.N, self.hamiltonian.address.M
        if _is_sparse(self.hamiltonian):  # the address of basis index i (or i itself when no addresses were given)
            names = self.hamiltonian.addresses
            for w in range(tau.shape[0]):
                for k in range(tau.shape[1]):
                    i = int(ad[w, k, 0])
                    yield dict(walker=w + 1, step=k + 1, address=i if names is None else names[i],
                               residence_time=tau[w, k], local_energy=el[w, k])
            return
        for w in range(tau.shape[0]):
            for k in range(tau.shape[1]):
                yield dict(walker=w + 1, step=k + 1, address=BoseFS.from_words(N, M, ad[w, k]),
                           residence_time=tau[w, k], local_energy=el[w, k])

    def curr_addresses(self) -> np.ndarray:
        return self.walkers_handle.addresses()

    def __repr__(self):
        try:
            est = local_energy_estimator(self)
            le = f"{est.mean:.5g} ± {est.err:.5g}"
        except Exception:  # pragma: no cover - cosmetic
            le = "n/a"
        return f"KineticVQMCResult\n  walkers:      {self.n_walkers}\n  samples:      {len(self)}\n  local energy: {le}"


def kinetic_vqmc(hamiltonian: HubbardReal1D, ansatz: AbstractAnsatz, params, *, samples=1e6, walkers=None,
                 steps=None, seed: int = DEFAULT_SEED, record=None, burn_in: int = 0,
                 kernel: int = _capi.KERNEL_BITPARALLEL, ctx: Context | None = None,
                 estimator: str = "walker_mean") -> KineticVQMCResult:
    """``kinetic_vqmc(hamiltonian, ansatz, params; samples=1e6, walkers, steps=round(Int, samples/walkers))``
    (qmc.jl:129-146).  Extra keywords: seed (Philox key), record (keep per-step series; default: when
    they fit in 256 MiB), burn_in (default 0 = reference behaviour), kernel, ctx, estimator ('walker_mean' =
    the reference's mean of per-walker ratio estimates; 'pooled' = one ratio over all walkers and steps, which
    has no per-walker O(tau_corr/steps) bias -- preferable for many short walkers; needs record=False)."""
    _check_pair(hamiltonian, ansatz)
    if walkers is None:
        walkers = default_walkers(samples)
    walkers = int(walkers)
    if steps is None:
        steps = int(round(samples / walkers))  # qmc.jl:133
    steps = int(steps)
    res = _make_result(hamiltonian, ansatz, params, walkers, steps, seed, record, burn_in, kernel, ctx, estimator)
    return res._run(steps, keep=False)


def _is_generic(ansatz) -> bool:
    """ansaetze served by the ansatz-generic engine (ansatz.py) rather than the Gutzwiller fast path"""
    return bool(getattr(ansatz, "generic", False))


def _is_sparse(hamiltonian) -> bool:
    """stored operators (sparse.py): any Hamiltonian materialised over its basis"""
    return bool(getattr(hamiltonian, "sparse_operator", False))


def _check_pair(hamiltonian, ansatz):
    if _is_sparse(hamiltonian) or getattr(ansatz, "sparse_ansatz", False):
        if not (_is_sparse(hamiltonian) and getattr(ansatz, "sparse_ansatz", False)):
            raise TypeError("a SparseHamiltonian needs an ansatz tabulated over its basis (sparse.py), and vice versa")
        if ansatz.hamiltonian is not hamiltonian:
            raise TypeError("the ansatz was built for a different Hamiltonian")
        return
    if getattr(ansatz, "vector_ansatz", False):
        return  # VectorAnsatz holds no Hamiltonian (vector.jl:6-8)
    if not isinstance(ansatz, AbstractAnsatz) or not hasattr(ansatz, "hamiltonian"):
        raise TypeError("ansatz must be one of the AbstractAnsatz types of this package")
    h2 = ansatz.hamiltonian
    if h2 is not hamiltonian and (h2.u, h2.t, getattr(h2, "v", 0.0), h2.address.N, h2.address.M) != \
            (hamiltonian.u, hamiltonian.t, getattr(hamiltonian, "v", 0.0), hamiltonian.address.N, hamiltonian.address.M):
        raise TypeError("the ansatz was built for a different Hamiltonian")
    if not _is_generic(ansatz) and not isinstance(ansatz, GutzwillerAnsatz):
        raise TypeError(f"unsupported ansatz {type(ansatz).__name__}")


def _make_result(hamiltonian, ansatz, params, walkers, steps, seed, record, burn_in, kernel, ctx,
                 estimator="walker_mean"):
    _check_pair(hamiltonian, ansatz)
    model = hamiltonian.model(ctx)
    words = 1 if _is_sparse(hamiltonian) else model.words
    world, rank = distributed_state() if (ctx is None) else (1, 0)
    begin, end = shard_range(walkers, world, rank)
    if end <= begin:
        raise ValueError("fewer walkers than GPUs")
    if estimator == "pooled":
        if record:
            raise ValueError("estimator='pooled' works on the running sums; use record=False")
        record = False
    if record is None:
        # like the reference, keep what was sampled: on the host while small (rows / DVec(res) need the addresses),
        # else the (tau, E_loc) series in HBM (enough for val_err_and_grad / local_energy_estimator), else nothing
        cells = (end - begin) * max(steps, 1)
        if cells * (16 + 8 * words) <= RECORD_LIMIT_BYTES and world == 1:
            record = True
        elif cells * 16 <= DEVICE_SERIES_LIMIT_BYTES and not _is_sparse(hamiltonian) and not _is_generic(ansatz) \
                and not getattr(ansatz, "vector_ansatz", False):
            record = "device"
        else:
            record = False
    if record not in (True, False, "device"):
        raise ValueError("record must be True (host series), 'device' (series kept in HBM) or False")
    if _is_sparse(hamiltonian):
        from .sparse import SparseWalkers
        w = SparseWalkers(model, ansatz, end - begin, hamiltonian.address.index, seed=seed, global_offset=begin)
    elif getattr(ansatz, "vector_ansatz", False):
        from .ansatz import VecWalkers
        w = VecWalkers(ansatz.handle(hamiltonian, ctx), end - begin, hamiltonian.address.words, seed=seed, global_offset=begin)
    elif _is_generic(ansatz):
        from .ansatz import GenWalkers
        w = GenWalkers(ansatz.handle(ctx), end - begin, hamiltonian.address.words, seed=seed, global_offset=begin)
    else:
        w = Walkers(model, end - begin, hamiltonian.address.words, seed=seed, global_offset=begin)
    if record == "device" and not isinstance(w, Walkers):
        raise ValueError("record='device' is available for HubbardReal1D + GutzwillerAnsatz walkers")
    return KineticVQMCResult(hamiltonian, ansatz, params, w, record, kernel, burn_in, estimator)


def kinetic_vqmc_continue(res: KineticVQMCResult, *, steps=None) -> KineticVQMCResult:
    """``kinetic_vqmc!(res; steps=length(res.states[1]))`` (qmc.jl:108-111): append `steps` more steps."""
    steps = res.steps if steps is None else int(steps)
    return res._run(steps, keep=True)


def _per_walker_estimates(res: KineticVQMCResult):
    return res.walkers_handle.estimates()


def val_and_grad_result(res: KineticVQMCResult):
    """state.jl:137-149: mean over walkers of the per-walker (val, grad) of state.jl:42-49."""
    r = res.last
    return r.val, np.array(r.grad[:], dtype=np.float64)


def val_err_and_grad_result(res: KineticVQMCResult):
    """state.jl:150-164: per-walker blocking of the resampled local energies, combined as sqrt(sum e_w^2)/walkers
    (Appendix B.6).  The series are analysed where they are: in HBM by the blocking kernel (record='device', and
    every ``val_err_and_grad(qmc, p)``), or on the host for small recorded runs.  Only a run that kept no series at
    all (record=False, other engines) falls back to the cross-walker standard error of the per-walker estimates."""
    r = res.last
    if res.has_device_series:  # device-side resample + blocking (gwz_blocking.cu)
        b = res.blocking_summary
        if b is None:
            b = res.blocking_summary = res.walkers_handle.blocking()
        return b.mean, b.err, np.array(b.grad[:], dtype=np.float64)
    if res.record is True and res._tau:
        tau, el = res.residence_times(), res.local_energies()
        gr = None
        vals = errs2 = 0.0
        grads = 0.0
        for w in range(tau.shape[0]):
            b = blocking_analysis(resample(tau[w], el[w]))
            vals += b.mean
            errs2 += b.err ** 2
            if gr is None:
                gr = _grad_ratio_series(res)
            grads = grads + 2.0 * ((tau[w] * (el[w] - b.mean)) @ gr[w]) / np.sum(tau[w])  # state.jl:57-58
        nw = tau.shape[0]
        return vals / nw, math.sqrt(errs2) / nw, np.atleast_1d(grads / nw)
    return r.val, (r.err if r.walkers > 1 else float("nan")), np.array(r.grad[:], dtype=np.float64)


def _grad_ratio_series(res: KineticVQMCResult) -> np.ndarray:
    """grad_ratios[walker, step, P] = grad psi / psi (qmc.jl:38; -diagonal_element for the Gutzwiller ansatz,
    gutzwiller.jl:30), recomputed on the GPU from the recorded addresses."""
    ad = res.addresses()
    nw, ns, W = ad.shape
    if _is_sparse(res.hamiltonian):  # the table itself holds grad psi / psi per basis index
        gr = np.asarray(res.ansatz.table(res.params)[1], dtype=np.float64).reshape(res.hamiltonian.dim, -1)
        return gr[ad[:, :, 0].astype(np.int64)]
    if getattr(res.ansatz, "vector_ansatz", False):
        return np.zeros((nw, ns, 0))
    if _is_generic(res.ansatz):
        out = res.walkers_handle.ansatz_handle.probe(ad.reshape(-1, W), res.params)
    else:
        out = res.hamiltonian.model(res.walkers_handle.model.ctx).probe(ad.reshape(-1, W), res.params)
    return out["grad_ratio"].reshape(nw, ns, -1)


def local_energy_estimator(res: KineticVQMCResult, resample_length=None) -> CombinedBlockingResult:
    """``local_energy_estimator(res; resample_length)`` (state.jl:265-285)."""
    if res.has_device_series:
        summ, a = res.walkers_handle.blocking(resample_length, per_walker=True)
        if summ.walkers != res.walkers_handle.n:  # sharded run: the per-walker rows of the other ranks are not here
            results = []
        elif res.walkers_handle.n <= 4096:
            results = [BlockingResult(a["mean"][i], a["err"][i], a["err_err"][i], a["err"][i] ** 2, int(a["k"][i]),
                                      int(a["blocks"][i])) for i in range(res.walkers_handle.n)]
        else:
            results = a  # dict of arrays: one Python object per walker would not scale to 2^24 walkers
        return CombinedBlockingResult(summ.mean, summ.err, summ.err_err, results)
    if res.record is True and res._tau:
        tau, el = res.residence_times(), res.local_energies()
        n_out = tau.shape[1] if resample_length is None else int(resample_length)
        results = [blocking_analysis(resample(tau[w], el[w], len=n_out)) for w in range(tau.shape[0])]
        n = len(results)
        return CombinedBlockingResult(
            float(np.mean([b.mean for b in results])),
            math.sqrt(float(np.mean([b.err ** 2 for b in results])) / n),
            math.sqrt(float(np.mean([b.err_err ** 2 for b in results])) / n), results)
    r = res.last
    return CombinedBlockingResult(r.val, r.err if r.walkers > 1 else float("nan"), float("nan"), [])


def mean(res: KineticVQMCResult) -> float:
    """``Statistics.mean(res)`` (state.jl:287-292)."""
    return res.last.val


def var(res: KineticVQMCResult) -> float:
    """``Statistics.var(res)`` (state.jl:293-298): mean over walkers of the time-weighted variance."""
    if res.record is True and res._tau:
        tau, el = res.residence_times(), res.local_energies()
        out = 0.0
        for w in range(tau.shape[0]):
            sw = np.sum(tau[w])
            m = np.sum(tau[w] * el[w]) / sw
            out += np.sum(tau[w] * (el[w] - m) ** 2) / sw  # StatsBase var(x, w): corrected = false by default
        return out / tau.shape[0]
    if hasattr(res.last, "walker_var"):
        return res.last.walker_var  # the same per-walker variances from the running sums, averaged on the device
    return res.last.pooled_var


def sampled_vector(res: KineticVQMCResult):
    """``DVec(res)`` (state.jl:303-319): deposit every recorded row's residence time on its address and
    normalise (2-norm).  Returns ``(addresses uint64[k, W], values float64[k])`` sorted by address."""
    tau, ad = res.residence_times(), res.addresses()
    W = ad.shape[-1]
    flat = np.ascontiguousarray(ad.reshape(-1, W))
    keys = flat.view([("w", np.uint64, (W,))]).ravel() if W > 1 else flat[:, 0]
    uniq, inverse = np.unique(keys, return_inverse=True)
    vals = np.bincount(inverse.ravel(), weights=tau.reshape(-1), minlength=len(uniq))
    vals = vals / np.linalg.norm(vals)  # normalize!(dv)
    addrs = uniq["w"] if W > 1 else uniq.reshape(-1, 1)
    return np.ascontiguousarray(addrs, dtype=np.uint64), vals


class KineticVQMC:
    """``KineticVQMC(hamiltonian, ansatz; samples=1e6, walkers, steps)`` (wrapper.jl:45-63).

    Walkers persist between calls (warm start, wrapper.jl:74-79): each call resets the estimators,
    keeps the positions, runs ``steps = samples ÷ walkers`` steps per walker and reduces.
    """

    def __init__(self, hamiltonian: HubbardReal1D, ansatz: AbstractAnsatz, *, samples=1e6, walkers=None, steps=None,
                 seed: int = DEFAULT_SEED, record=None, burn_in: int = 0, kernel: int = _capi.KERNEL_BITPARALLEL,
                 ctx: Context | None = None, estimator: str = "walker_mean"):
        if walkers is None:
            walkers = default_walkers(samples)
        self.hamiltonian, self.ansatz = hamiltonian, ansatz
        self.walkers = int(walkers)
        self.steps = int(round(samples / self.walkers)) if steps is None else int(steps)  # wrapper.jl:56
        if record is None:
            record = False  # val_err_and_grad keeps the series in HBM for its own call; nothing is shipped to the host
        self.result = _make_result(hamiltonian, ansatz, np.zeros(ansatz.N_PARAMS), self.walkers, self.steps, seed,
                                   record, burn_in, kernel, ctx, estimator)

    def _reset_and_run(self, params, samples, steps, want_err=False):
        if steps is None:
            steps = (self.steps * self.walkers if samples is None else int(samples)) // self.walkers  # wrapper.jl:83
        res = self.result
        res.params = as_params(params, self.ansatz.N_PARAMS)  # _reset! (wrapper.jl:74-79)
        return res._run(int(steps), keep=False, want_err=want_err)

    def __call__(self, *args, samples=None, steps=None):
        if len(args) == 3:  # qmc(F, G, params): Optim.only_fg! protocol (wrapper.jl:107-121)
            F, G, params = args
            res = self._reset_and_run(params, None, self.steps)
            if G is not None:
                val, grad = val_and_grad_result(res)
                G[...] = grad
            else:
                val = mean(res)
            return val if F is not None else None
        (params,) = args
        return mean(self._reset_and_run(params, samples, steps))  # wrapper.jl:81-88

    def val_and_grad(self, params, *, samples=None, steps=None):
        return val_and_grad_result(self._reset_and_run(params, samples, steps))  # wrapper.jl:90-97

    def val_err_and_grad(self, params, *, samples=None, steps=None):
        return val_err_and_grad_result(self._reset_and_run(params, samples, steps, want_err=True))  # wrapper.jl:98-105

    def __repr__(self):
        return (f"KineticVQMC(\n  {self.hamiltonian!r},\n  {self.ansatz!r};\n  steps={self.steps},\n"
                f"  walkers={self.walkers},\n)")
