"""LocalEnergyEvaluator -- host mirror of src/localenergy.jl:81-173 over the CUDA sweep kernel."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from ._capi import check, lib
from .hamiltonian import AbstractAnsatz, BoseFS, GutzwillerAnsatz, HubbardReal1D, as_params
from .runtime import Context, distributed_state, shard_range


class LocalEnergyEvaluator:
    """``LocalEnergyEvaluator(hamiltonian, ansatz; basis=build_basis(hamiltonian))`` (localenergy.jl:81-88).

    ``basis`` may be a sequence of :class:`BoseFS` or a ``uint64[dim, W]`` array (uploaded to HBM once,
    the reference's ``basis::Vector{A}``).  Without a basis all C(N+M-1, N) Fock states are enumerated
    on the GPU by rank (needs N+M-1 <= 64); across several GPUs each rank sweeps its share of the
    rank range and the four sums are all-reduced.
    """

    def __init__(self, hamiltonian: HubbardReal1D, ansatz: AbstractAnsatz, *, basis=None, ctx: Context | None = None):
        self._gen = None
        self._vec = None
        self._sp = None
        if getattr(hamiltonian, "sparse_operator", False):  # stored operator (sparse.py): basis = its rows
            if not getattr(ansatz, "sparse_ansatz", False) or ansatz.hamiltonian is not hamiltonian:
                raise TypeError("a SparseHamiltonian needs an ansatz tabulated over its basis")
            if basis is not None:
                raise ValueError("a SparseHamiltonian is already materialised over its basis; do not pass one")
            self.hamiltonian, self.ansatz, self.P = hamiltonian, ansatz, ansatz.N_PARAMS
            self.model = self._sp = hamiltonian.handle(ctx)
            self._h = self._sp.handle
            world, rank = distributed_state() if ctx is None else (1, 0)
            self._rows = shard_range(hamiltonian.dim, world, rank)
            return
        if getattr(ansatz, "vector_ansatz", False):  # basis = keys(vector) (vector.jl:10): the Rayleigh quotient
            if basis is not None:
                raise ValueError("LocalEnergyEvaluator(H, VectorAnsatz(v)) sweeps keys(v); do not pass a basis")
            self.hamiltonian, self.ansatz, self.model, self.P = hamiltonian, ansatz, hamiltonian.model(ctx), 0
            self._vec = ansatz.handle(hamiltonian, ctx)
            self._h = self._vec.handle
            return
        self.hamiltonian, self.ansatz = hamiltonian, ansatz
        self.model = hamiltonian.model(ctx)
        self.P = ansatz.N_PARAMS
        if getattr(ansatz, "generic", False):  # the ansatz-generic engine (ansatz.py)
            from .ansatz import GenSweep
            arr = None
            begin = end = 0
            if basis is not None:
                if len(basis) and isinstance(basis[0], BoseFS):
                    basis = np.stack([b.words for b in basis])
                arr = np.ascontiguousarray(basis, dtype=np.uint64).reshape(-1, self.model.words)
                if arr.shape[0] == 0:
                    raise ValueError("empty basis")
            else:
                world, rank = distributed_state() if ctx is None else (1, 0)
                begin, end = shard_range(self.model.dimension(), world, rank)
            self._gen = GenSweep(ansatz.handle(ctx), arr, begin, end)
            self._h = self._gen.handle
            return
        if not isinstance(ansatz, GutzwillerAnsatz):
            raise TypeError(f"unsupported ansatz {type(ansatz).__name__}")
        self._h = C.c_void_p()
        if basis is not None:
            if len(basis) and isinstance(basis[0], BoseFS):
                basis = np.stack([b.words for b in basis])
            arr = np.ascontiguousarray(basis, dtype=np.uint64).reshape(-1, self.model.words)
            if arr.shape[0] == 0:
                raise ValueError("empty basis")
            check(lib().gwz_sweep_create(self.model.handle, arr.ctypes.data_as(C.c_void_p), arr.shape[0], 0, 0,
                                         C.byref(self._h)))
        else:
            world, rank = distributed_state() if ctx is None else (1, 0)
            begin, end = shard_range(self.model.dimension(), world, rank)
            check(lib().gwz_sweep_create(self.model.handle, None, 0, begin, end, C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_sweep_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def is_bit_level(self) -> bool:
        """True when ``handle`` is a ``gwz_sweep`` (HubbardReal1D + GutzwillerAnsatz kernels), i.e. none of the other
        engines' handles (generic ansatz, stored operator, vector ansatz)."""
        return self._sp is None and self._gen is None and self._vec is None

    @property
    def dim(self) -> int:
        if self._sp is not None:
            return self.hamiltonian.dim
        if self._vec is not None:
            return int(self.ansatz.values.size)
        if self._gen is not None:
            return self._gen.dim()
        d = C.c_uint64()
        check(lib().gwz_sweep_dim(self._h, C.byref(d)))
        return d.value

    def val_and_grad(self, params):
        """``val_and_grad(le, params)`` (localenergy.jl:94-129)."""
        if self._sp is not None:
            self._sp.load(self.ansatz, params)
            return self._sp.sweep(True, *self._rows)
        if self._vec is not None:
            return self._vec.rayleigh(), np.zeros(0)
        if self._gen is not None:
            return self._gen.val_and_grad(params)
        p = as_params(params, 1)
        val, grad = C.c_double(), (C.c_double * 1)()
        check(lib().gwz_sweep_val_and_grad(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(val), grad, None))
        return val.value, np.array([grad[0]])

    def accumulators(self, params) -> np.ndarray:
        """(numerator, denominator, numerator_grad, denominator_grad) of localenergy.jl:121."""
        p = as_params(params, 1)
        val, grad, acc = C.c_double(), (C.c_double * 1)(), (C.c_double * 4)()
        check(lib().gwz_sweep_val_and_grad(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(val), grad, acc))
        return np.array(acc[:])

    def __call__(self, *args):
        if len(args) == 3 and (self.P != 3 or args[0] is None or args[1] is None or not np.isscalar(args[0])):
            F, G, params = args  # le(F, G, params): Optim.only_fg! protocol (localenergy.jl:157-169)
            if G is not None:
                val, grad = self.val_and_grad(params)
                G[...] = grad
            else:
                val = self(params)
            return val if F is not None else None
        params = args[0] if len(args) == 1 else args  # le(p...) (localenergy.jl:171-173)
        if self._sp is not None:
            self._sp.load(self.ansatz, params)
            return self._sp.sweep(False, *self._rows)[0]
        if self._vec is not None:
            return self._vec.rayleigh()
        if self._gen is not None:
            return self._gen.val(params)
        p = as_params(params, 1)
        val = C.c_double()
        check(lib().gwz_sweep_val(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(val)))
        return val.value  # localenergy.jl:131-155

    def terms(self, params, first: int = 0, count: int | None = None) -> np.ndarray:
        """Per-state (num, den, num_grad[P], den_grad[P]) of localenergy.jl:105-120: float64[count, 2 + 2P]."""
        count = self.dim - first if count is None else int(count)
        if self._sp is not None:
            self._sp.load(self.ansatz, params)
            return self._sp.sweep_terms(first, count)
        if self._gen is not None:
            return self._gen.terms(params, first, count)
        p = as_params(params, 1)
        out = np.zeros((count, 4))
        check(lib().gwz_sweep_terms(self._h, p.ctypes.data_as(C.POINTER(C.c_double)), int(first), count,
                                    out.ctypes.data_as(C.c_void_p)))
        return out

    def states(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.dim - first if count is None else int(count)
        if self._sp is not None:
            return np.arange(first, first + count, dtype=np.uint64).reshape(-1, 1)  # basis indices
        if self._gen is not None:
            return self._gen.states(first, count)
        out = np.zeros((count, self.model.words), dtype=np.uint64)
        check(lib().gwz_sweep_states(self._h, int(first), count, out.ctypes.data_as(C.c_void_p)))
        return out

    def last_kernel_ms(self) -> float:
        if self._sp is not None:
            return self._sp.last_kernel_ms()
        if self._gen is not None:
            return self._gen.last_kernel_ms()
        ms = C.c_float()
        check(lib().gwz_sweep_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def __repr__(self):
        return f"LocalEnergyEvaluator({self.hamiltonian!r}, {self.ansatz!r})"
