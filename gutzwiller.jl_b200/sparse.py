"""Stored operators: kinetic VQMC and LocalEnergyEvaluator for ANY Hamiltonian and ANY ansatz.

The reference's hot path is generic in the Hamiltonian -- ``kinetic_sample!`` (qmc.jl:14-44) and
``LocalEnergyEvaluator`` (localenergy.jl:94-155) only call ``offdiagonals``, ``diagonal_element`` and the ansatz --
and its tests run it on ``HubbardMom1D`` and a two-component ``HubbardRealSpace`` as well (test/qmc.jl:10-14).  Such
operators reach the GPU materialised over ``build_basis(H)``: :class:`SparseHamiltonian` holds the rows
(``offdiagonals(H, basis[k])`` in the reference's order), and an ansatz is a table over the basis
(:class:`SparseGutzwillerAnsatz` is filled on the device, :class:`TableAnsatz` / :class:`SparseVectorAnsatz` are
uploaded).  ``kinetic_vqmc``, ``KineticVQMC``, ``LocalEnergyEvaluator`` and ``amsgrad`` accept these objects in place
of ``HubbardReal1D`` + its ansaetze; walker "addresses" are basis indices.  All arithmetic runs in
csrc/gwz_sparse.cu behind the ``gwz_sparse_*`` / ``gwz_swalkers_*`` C ABI.
"""
from __future__ import annotations

import ctypes as C
import itertools
import weakref

import numpy as np

from . import _capi
from ._capi import check, lib
from .hamiltonian import AbstractAnsatz, as_params
from .runtime import Context, default_context


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_ansatz_tokens = itertools.count(1)  # identity of an ansatz object for the table cache (id() values are reused)


class _SparseHandle:
    """gwz_sparse on one context, with the identity of the table it currently holds"""

    def __init__(self, H: "SparseHamiltonian", ctx: Context):
        self.ctx, self.H = ctx, H
        self._h = C.c_void_p()
        check(lib().gwz_sparse_create(ctx.handle, H.dim, _vp(H.row_ptr), _vp(H.col), _vp(H.melem), _vp(H.diag),
                                      C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_sparse_destroy, self._h)
        self.loaded = None  # (ansatz id, params bytes)

    @property
    def handle(self):
        return self._h

    def info(self) -> dict:
        d, nz, mr = C.c_uint64(), C.c_uint64(), C.c_uint64()
        P, G = C.c_int32(), C.c_int32()
        check(lib().gwz_sparse_info(self._h, C.byref(d), C.byref(nz), C.byref(mr), C.byref(P), C.byref(G)))
        rb, bb = C.c_int32(), C.c_uint64()
        check(lib().gwz_sparse_layout(self._h, C.byref(rb), C.byref(bb)))
        return dict(dim=d.value, nnz=nz.value, max_row=mr.value, num_parameters=P.value, lanes_per_walker=G.value,
                    row_blocks=bool(rb.value), block_bytes=bb.value)

    def set_table(self, psi, grad_ratio=None) -> None:
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        if psi.shape != (self.H.dim,):
            raise ValueError(f"amplitude table must have {self.H.dim} entries")
        P = 0
        if grad_ratio is not None:
            grad_ratio = np.ascontiguousarray(grad_ratio, dtype=np.float64).reshape(self.H.dim, -1)
            P = grad_ratio.shape[1]
        check(lib().gwz_sparse_set_table(self._h, _vp(psi), _vp(grad_ratio) if P else None, P))
        self.loaded = None

    def set_gutzwiller(self, g: float) -> None:
        check(lib().gwz_sparse_set_gutzwiller(self._h, float(g)))
        self.loaded = None

    def get_table(self):
        P = self.info()["num_parameters"]
        psi, gr = np.zeros(self.H.dim), np.zeros((self.H.dim, P))
        check(lib().gwz_sparse_get_table(self._h, _vp(psi), _vp(gr) if P else None))
        return psi, gr

    def load(self, ansatz, params) -> None:
        """make the device table that of ansatz(., params) (cached: the same (ansatz, params) is not re-uploaded)"""
        p = as_params(params if np.size(params) else np.zeros(0), ansatz.N_PARAMS)
        key = (ansatz._token, p.tobytes())
        if self.loaded == key:
            return
        ansatz._load(self, p)
        self.loaded = key

    def probe(self, idx, u01=None, want_rates: bool = False) -> dict:
        """kinetic_sample! per index (qmc.jl:14-44) with the table currently loaded"""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        n, info = idx.shape[0], self.info()
        P, mr = info["num_parameters"], max(info["max_row"], 1)
        u = None if u01 is None else np.ascontiguousarray(u01, dtype=np.float64)
        out = dict(e_loc=np.zeros(n), residence_time=np.zeros(n), grad_ratio=np.zeros((n, P)),
                   n_hops=np.zeros(n, dtype=np.int32), chosen=np.zeros(n, dtype=np.int32),
                   next_index=np.zeros(n, dtype=np.uint32), rates=np.zeros((n, mr)) if want_rates else None)
        check(lib().gwz_sparse_probe(self._h, _vp(idx), n, _vp(u), _vp(out["e_loc"]), _vp(out["residence_time"]),
                                     _vp(out["grad_ratio"]) if P else None, _vp(out["n_hops"]), _vp(out["rates"]),
                                     _vp(out["chosen"]), _vp(out["next_index"])))
        return out

    def sweep(self, want_grad: bool = True, row_begin: int = 0, row_end: int | None = None):
        P = self.info()["num_parameters"]
        val, grad = C.c_double(), np.zeros(max(P, 1))
        end = self.H.dim if row_end is None else int(row_end)
        check(lib().gwz_sparse_sweep(self._h, int(row_begin), end, C.byref(val), _vp(grad) if want_grad and P else None))
        return val.value, grad[:P].copy()

    def sweep_terms(self, first: int = 0, count: int | None = None) -> np.ndarray:
        P = self.info()["num_parameters"]
        count = self.H.dim - first if count is None else int(count)
        out = np.zeros((count, 2 + 2 * P))
        check(lib().gwz_sparse_sweep_terms(self._h, int(first), count, _vp(out)))
        return out

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(lib().gwz_sparse_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class _IndexAddress:
    """stands where ``hamiltonian.address`` is expected: a basis index is one 64-bit word"""
    N = M = 0

    def __init__(self, index: int):
        self.index = int(index)
        self.words = np.array([self.index], dtype=np.uint64)


class SparseHamiltonian:
    """Any Hamiltonian materialised over its basis.

    ``row_ptr[dim + 1]``, ``col[nnz]``, ``melem[nnz]``: row k = ``offdiagonals(H, basis[k])`` in the reference's
    order; ``diag[dim]`` = ``diagonal_element(H, basis[k])``; ``addresses`` (optional) = ``build_basis(H)`` for
    reporting; ``start`` = index of ``starting_address(H)`` (0 for a breadth-first ``build_basis``).
    The Julia shim builds these arrays from any Rimu Hamiltonian (julia/GutzwillerB200.jl, ``GPUSparseHamiltonian(H)``)."""

    sparse_operator = True

    def __init__(self, row_ptr, col, melem, diag, addresses=None, start: int = 0):
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.melem = np.ascontiguousarray(melem, dtype=np.float64)
        self.diag = np.ascontiguousarray(diag, dtype=np.float64)
        self.dim = int(self.diag.shape[0])
        if self.row_ptr.shape != (self.dim + 1,):
            raise ValueError("row_ptr must have dim + 1 entries")
        if self.col.shape != self.melem.shape or self.col.shape != (int(self.row_ptr[-1]),):
            raise ValueError("col and melem must have row_ptr[-1] entries")
        if not 0 <= int(start) < self.dim:
            raise ValueError("starting index outside the basis")
        self.addresses = addresses
        self.address = _IndexAddress(start)
        self._handles = {}

    @classmethod
    def from_hubbard(cls, H, max_dim: int = 5_000_000) -> "SparseHamiltonian":
        """Materialise ``HubbardReal1D`` / ``ExtendedHubbardReal1D`` over ALL C(N+M-1, N) Fock states (host-side
        numpy): what ``build_basis`` + ``offdiagonals`` + ``diagonal_element`` give, rows in
        the reference's hop order (occupied modes ascending, right hop then left hop).  Basis index = colex rank of
        the unary bitstring, i.e. the order of the ranked ``LocalEnergyEvaluator`` sweep; ``addresses`` holds the
        occupation numbers.  This is the route for a user-defined ansatz on the chain: tabulate it over
        ``Hs.addresses`` (:class:`TableAnsatz`) and every call form works unchanged."""
        from math import comb
        N, M = H.address.N, H.address.M
        dim = comb(N + M - 1, N)
        if dim > max_dim:
            raise ValueError(f"BoseFS{{{N},{M}}} has {dim} states: too many to store (max_dim={max_dim})")
        B = N + M - 1
        binom = np.array([[comb(p, i) for i in range(N + 1)] for p in range(B + 1)], dtype=np.int64)

        def rank(onr):  # colex rank of the positions of the ones: boson j (sorted by mode) sits at bit mode_j + j
            modes = np.repeat(np.tile(np.arange(M), onr.shape[0]), onr.ravel()).reshape(onr.shape[0], N)
            return binom[modes + np.arange(N), np.arange(1, N + 1)].sum(axis=1)

        # all compositions of N into M parts, then sorted by rank
        onr = np.zeros((1, M), dtype=np.int64)
        onr[0, 0] = N
        parts = [np.array([[n]], dtype=np.int64) for n in range(N + 1)]  # compositions of n into 1 part
        for _ in range(M - 1):
            nxt = []
            for n in range(N + 1):
                nxt.append(np.concatenate([np.hstack([parts[n - k], np.full((parts[n - k].shape[0], 1), k)])
                                           for k in range(n + 1)]))
            parts = nxt
        onr = parts[N]
        order = np.argsort(rank(onr), kind="stable")
        onr = np.ascontiguousarray(onr[order])
        assert onr.shape[0] == dim and np.array_equal(rank(onr), np.arange(dim))
        u, t, v = float(H.u), float(H.t), float(getattr(H, "v", 0.0))
        diag = 0.5 * u * (onr * (onr - 1)).sum(axis=1) + v * (onr * np.roll(onr, -1, axis=1)).sum(axis=1)
        target = np.zeros((dim, M, 2), dtype=np.int64)
        melem = np.zeros((dim, M, 2))
        for i in range(M):
            for d, j in ((0, (i + 1) % M), (1, (i - 1) % M)):  # right hop, then left hop (App. A.4)
                src = np.flatnonzero(onr[:, i] > 0)
                new = onr[src].copy()
                melem[src, i, d] = -t * np.sqrt(new[:, i] * (new[:, j] + 1.0))
                new[:, i] -= 1
                new[:, j] += 1
                target[src, i, d] = rank(new)
        valid = np.repeat((onr > 0)[:, :, None], 2, axis=2)
        row_ptr = np.zeros(dim + 1, dtype=np.uint64)
        row_ptr[1:] = np.cumsum(2 * (onr > 0).sum(axis=1))
        start = int(rank(np.array([H.address.onr()], dtype=np.int64))[0])
        return cls(row_ptr, target[valid].astype(np.uint32), melem[valid], diag, addresses=onr, start=start)

    def handle(self, ctx: Context | None = None) -> _SparseHandle:
        ctx = ctx or default_context()
        h = self._handles.get(id(ctx))
        if h is None:
            h = self._handles[id(ctx)] = _SparseHandle(self, ctx)
        return h

    model = handle  # what vqmc.py asks a Hamiltonian for

    def __repr__(self):
        return f"SparseHamiltonian(dim={self.dim}, nnz={self.col.size})"


class SparseAnsatz(AbstractAnsatz):
    """An ansatz tabulated over the basis of a :class:`SparseHamiltonian`.  Subclasses give ``table(params) ->
    (psi[dim], grad_ratio[dim, P])``."""

    sparse_ansatz = True
    N_PARAMS = 0

    def __init__(self, hamiltonian: SparseHamiltonian):
        if not getattr(hamiltonian, "sparse_operator", False):
            raise TypeError("a SparseHamiltonian is required")
        self.hamiltonian = hamiltonian
        self._token = next(_ansatz_tokens)

    def table(self, params):
        raise NotImplementedError

    def _load(self, handle: _SparseHandle, p: np.ndarray) -> None:
        psi, gr = self.table(p)
        handle.set_table(psi, gr if self.N_PARAMS else None)

    def __call__(self, index: int, params=()) -> float:
        return float(self.table(as_params(params if np.size(params) else np.zeros(0), self.N_PARAMS))[0][int(index)])

    def val_and_grad(self, index: int, params=()):
        psi, gr = self.table(as_params(params if np.size(params) else np.zeros(0), self.N_PARAMS))
        v = float(psi[int(index)])
        return v, (np.asarray(gr).reshape(self.hamiltonian.dim, -1)[int(index)] * v if self.N_PARAMS else np.zeros(0))


class SparseGutzwillerAnsatz(SparseAnsatz):
    """``GutzwillerAnsatz(H)`` (gutzwiller.jl:14-36) for a stored operator: psi_k = exp(-g diag_k); the table is
    filled on the device."""

    N_PARAMS = 1

    def table(self, params):
        g = float(as_params(params, 1)[0])
        d = self.hamiltonian.diag
        return np.exp(-g * d), -d.reshape(-1, 1)

    def _load(self, handle, p):
        handle.set_gutzwiller(float(p[0]))

    def __repr__(self):
        return f"GutzwillerAnsatz({self.hamiltonian!r})"


class SparseVectorAnsatz(SparseAnsatz):
    """``VectorAnsatz(vector)`` (vector.jl:6-12) over the basis of a stored operator: 0 parameters."""

    N_PARAMS = 0

    def __init__(self, hamiltonian: SparseHamiltonian, values):
        super().__init__(hamiltonian)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        if self.values.shape != (hamiltonian.dim,):
            raise ValueError("one amplitude per basis state is required")

    def table(self, params=()):
        return self.values, np.zeros((self.hamiltonian.dim, 0))

    def __repr__(self):
        return f"VectorAnsatz({self.values.size} entries)"


class TableAnsatz(SparseAnsatz):
    """Any ansatz evaluated by the caller: ``fn(params) -> (psi[dim], grad_ratio[dim, P])`` with
    grad_ratio = (d psi / d p) / psi -- e.g. the reference's Jastrow-type ansaetze on HubbardMom1D
    (test/ansatz.jl:44-60), evaluated over build_basis on the host."""

    def __init__(self, hamiltonian: SparseHamiltonian, fn, num_parameters: int):
        super().__init__(hamiltonian)
        self.fn, self.N_PARAMS = fn, int(num_parameters)

    def table(self, params):
        psi, gr = self.fn(as_params(params if np.size(params) else np.zeros(0), self.N_PARAMS))
        return np.asarray(psi, dtype=np.float64), np.asarray(gr, dtype=np.float64).reshape(self.hamiltonian.dim, -1)

    def __repr__(self):
        return f"TableAnsatz({self.hamiltonian!r}, P={self.N_PARAMS})"


class SparseRunResult:
    """attribute names follow ``gwz_vqmc_result`` (as GenRunResult in ansatz.py)"""

    def __init__(self, r: _capi.GenResult, grad, grad_err, pooled_grad):
        self.val, self.err, self.pooled_val, self.pooled_var = r.val, r.err, r.pooled_val, r.pooled_var
        self.total_time, self.walkers, self.steps, self.hops = r.total_time, r.walkers, r.steps, r.hops
        self.kernel_ms = r.kernel_ms
        self.grad, self.grad_err, self.pooled_grad = grad, grad_err, pooled_grad


class SparseWalkers:
    """gwz_swalkers: n walkers on a stored operator (global ids offset..offset+n-1) resident on one device"""

    def __init__(self, shandle: _SparseHandle, ansatz: SparseAnsatz, n_walkers: int, start_index: int, seed: int,
                 global_offset: int = 0):
        self.shandle, self.ansatz, self.model = shandle, ansatz, shandle
        self.n, self.offset = int(n_walkers), int(global_offset)
        self._h = C.c_void_p()
        check(lib().gwz_swalkers_create(shandle.handle, self.n, self.offset, int(start_index),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_swalkers_destroy, self._h)

    @property
    def handle(self):
        return self._h

    @property
    def P(self) -> int:
        return self.ansatz.N_PARAMS

    def indices(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint32)
        check(lib().gwz_swalkers_get_indices(self._h, _vp(out)))
        return out

    def set_indices(self, idx) -> None:
        a = np.ascontiguousarray(idx, dtype=np.uint32).reshape(self.n)
        check(lib().gwz_swalkers_set_indices(self._h, _vp(a)))

    def addresses(self) -> np.ndarray:
        return self.indices().astype(np.uint64).reshape(-1, 1)

    def set_step_counter(self, steps_done: int) -> None:
        check(lib().gwz_swalkers_set_step_counter(self._h, int(steps_done)))

    def _flags(self, keep_acc, pooled=False):
        return (_capi.RUN_KEEP_ACC if keep_acc else 0) | (_capi.RUN_POOLED if pooled else 0)

    def run(self, params, steps: int, burn_in: int = 0, kernel=None, keep_acc: bool = False, pooled: bool = False):
        self.shandle.load(self.ansatz, params)
        P = self.P
        r = _capi.GenResult()
        g, ge, pg = np.zeros(max(P, 1)), np.zeros(max(P, 1)), np.zeros(max(P, 1))
        check(lib().gwz_swalkers_run(self._h, int(steps), int(burn_in), self._flags(keep_acc, pooled), C.byref(r),
                                     _vp(g), _vp(ge), _vp(pg)))
        return SparseRunResult(r, g[:P], ge[:P], pg[:P])

    def run_record(self, params, steps: int, kernel=None, keep_acc: bool = False, want_addresses: bool = True):
        self.shandle.load(self.ansatz, params)
        P, steps = self.P, int(steps)
        tau, eloc = np.zeros((steps, self.n)), np.zeros((steps, self.n))
        idx = np.zeros((steps, self.n), dtype=np.uint32) if want_addresses else None
        r = _capi.GenResult()
        g, ge, pg = np.zeros(max(P, 1)), np.zeros(max(P, 1)), np.zeros(max(P, 1))
        check(lib().gwz_swalkers_run_record(self._h, steps, self._flags(keep_acc), _vp(tau), _vp(eloc), _vp(idx),
                                            C.byref(r), _vp(g), _vp(ge), _vp(pg)))
        addr = None if idx is None else idx.astype(np.uint64).reshape(steps, self.n, 1)
        return SparseRunResult(r, g[:P], ge[:P], pg[:P]), tau, eloc, addr

    def sample_vector(self, params, steps: int, burn_in: int = 0, keep_acc: bool = False):
        """``DVec(res)`` (state.jl:303-319) accumulated on the device: total residence time per basis index"""
        self.shandle.load(self.ansatz, params)
        hist = np.zeros(self.shandle.H.dim)
        r = _capi.GenResult()
        check(lib().gwz_swalkers_sample_vector(self._h, int(steps), int(burn_in), self._flags(keep_acc), _vp(hist),
                                               C.byref(r)))
        z = np.zeros(0)
        return hist, SparseRunResult(r, z, z, z)

    def estimates(self):
        v, g = np.zeros(self.n), np.zeros((self.n, max(self.P, 1)))
        check(lib().gwz_swalkers_estimates(self._h, _vp(v), _vp(g)))
        return v, g[:, :self.P]
