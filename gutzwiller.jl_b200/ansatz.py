"""The reference's other single-component ansaetze and ExtendedHubbardReal1D -- host mirror over the
ansatz-generic CUDA engine (csrc/gwz_generic.cu, C ABI ``gwz_gen_*`` / ``gwz_gwalkers_*`` / ``gwz_gsweep_*``).

    ExtendedGutzwillerAnsatz   src/ansatz/extgutzwiller.jl:14-35
    MultinomialAnsatz          src/ansatz/multinomial.jl:17-59
    DensityProfileAnsatz       src/ansatz/densityprofile.jl:8-28
    RelativeJastrowAnsatz      src/ansatz/jastrow.jl:71-103
    JastrowAnsatz              src/ansatz/jastrow.jl:13-58
    GutzwillerAnsatz(ExtendedHubbardReal1D)   src/ansatz/gutzwiller.jl:14-36 with Rimu's extended Hamiltonian

They plug into the same ``kinetic_vqmc`` / ``KineticVQMC`` / ``LocalEnergyEvaluator`` / ``amsgrad`` calls as
``GutzwillerAnsatz(HubbardReal1D)``; the objects only hold parameters and handles, all arithmetic is on the GPU.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _capi
from ._capi import check, lib
from .hamiltonian import AbstractAnsatz, BoseFS, GutzwillerAnsatz, HubbardReal1D, as_params
from .runtime import Context, default_context

_DP = C.POINTER(C.c_double)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ExtendedHubbardReal1D(HubbardReal1D):
    """Rimu ``ExtendedHubbardReal1D(addr; u=1.0, v=1.0, t=1.0)``: HubbardReal1D plus the nearest-neighbour
    interaction ``v * sum_i n_i n_{i+1}`` (periodic)."""

    def __init__(self, addr: BoseFS, u: float = 1.0, v: float = 1.0, t: float = 1.0):
        super().__init__(addr, u=u, t=t)
        self.v = float(v)

    def __repr__(self):
        return f"ExtendedHubbardReal1D({self.address!r}; u={self.u}, v={self.v}, t={self.t})"


class _AnsatzHandle:
    """gwz_ansatz on one model (one context)."""

    def __init__(self, model, kind, normalize):
        """kind / normalize: ints, or pairs for a CombinationAnsatz (a1 + a2)."""
        self.model = model
        self._h = C.c_void_p()
        if isinstance(kind, tuple):
            check(lib().gwz_ansatz_create_combination(model.handle, int(kind[0]), int(bool(normalize[0])), int(kind[1]),
                                                      int(bool(normalize[1])), C.byref(self._h)))
        else:
            check(lib().gwz_ansatz_create(model.handle, int(kind), int(bool(normalize)), C.byref(self._h)))
        k, p = C.c_int(), C.c_int()
        check(lib().gwz_ansatz_info(self._h, C.byref(k), C.byref(p)))
        self.kind, self.P = k.value, p.value
        self._finalizer = weakref.finalize(self, lib().gwz_ansatz_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def eval(self, addrs, params, want_grad=True):
        addrs = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(-1, self.model.words)
        n = addrs.shape[0]
        p = as_params(params, self.P)
        val = np.zeros(n)
        grad = np.zeros((n, self.P)) if want_grad else None
        check(lib().gwz_gen_ansatz_eval(self._h, _vp(addrs), n, p.ctypes.data_as(_DP), _vp(val), _vp(grad)))
        return val, grad

    def probe(self, addrs, params, u01=None, want_rates=False) -> dict:
        """One kinetic_sample! per address (qmc.jl:14-44): same keys as :meth:`runtime.Model.probe`."""
        addrs = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(-1, self.model.words)
        n = addrs.shape[0]
        p = as_params(params, self.P)
        u = None if u01 is None else np.ascontiguousarray(u01, dtype=np.float64)
        if u is not None and u.shape != (n,):
            raise ValueError("u01 must have one entry per address")
        out = dict(e_loc=np.zeros(n), residence_time=np.zeros(n), grad_ratio=np.zeros((n, self.P)),
                   n_hops=np.zeros(n, dtype=np.int32), chosen=np.zeros(n, dtype=np.int32),
                   next_addr=np.zeros((n, self.model.words), dtype=np.uint64))
        rates = np.zeros((n, 2 * self.model.M)) if want_rates else None
        check(lib().gwz_gen_probe_addresses(self._h, _vp(addrs), n, p.ctypes.data_as(_DP), _vp(u), _vp(out["e_loc"]),
                                            _vp(out["residence_time"]), _vp(out["grad_ratio"]), _vp(out["n_hops"]),
                                            _vp(rates), _vp(out["chosen"]), _vp(out["next_addr"])))
        if want_rates:
            out["rates"] = rates
        return out


class GenericAnsatz(AbstractAnsatz):
    """Base of the ansaetze served by the generic engine.  ``ansatz(addr, params)`` and
    ``val_and_grad(ansatz, addr, params)`` follow src/ansatz/abstract.jl:15-22,50-52."""
    KIND = -1
    generic = True

    def __init__(self, hamiltonian: HubbardReal1D, normalize: bool = False):
        if not isinstance(hamiltonian, HubbardReal1D):
            raise TypeError("needs a HubbardReal1D / ExtendedHubbardReal1D Hamiltonian")
        self.hamiltonian = hamiltonian
        self.normalize = bool(normalize)
        self.N_PARAMS = self._num_parameters(hamiltonian.address.M)
        self._handles: dict[int, _AnsatzHandle] = {}

    @staticmethod
    def _num_parameters(M: int) -> int:
        raise NotImplementedError

    def handle(self, ctx: Context | None = None) -> _AnsatzHandle:
        ctx = ctx or default_context()
        key = id(ctx)
        if key not in self._handles:
            self._handles[key] = _AnsatzHandle(self.hamiltonian.model(ctx), self.KIND, self.normalize)
        return self._handles[key]

    def __call__(self, addr: BoseFS, params) -> float:
        return float(self.handle().eval(addr.words[None, :], params, want_grad=False)[0][0])

    def val_and_grad(self, addr: BoseFS, params):
        val, grad = self.handle().eval(addr.words[None, :], params)
        return float(val[0]), grad[0].copy()

    def __repr__(self):
        return f"{type(self).__name__}({self.hamiltonian!r})"


class GenericGutzwillerAnsatz(GenericAnsatz, GutzwillerAnsatz):
    """``GutzwillerAnsatz(H)`` with ``H::ExtendedHubbardReal1D``: exp(-g <f|H|f>) with the extended diagonal."""
    KIND = _capi.ANSATZ_GUTZWILLER

    def __new__(cls, hamiltonian=None):
        return object.__new__(cls)

    @staticmethod
    def _num_parameters(M):
        return 1

    def __repr__(self):
        return f"GutzwillerAnsatz({self.hamiltonian!r})"


class ExtendedGutzwillerAnsatz(GenericAnsatz):
    """``exp(-g1 sum n_i(n_i-1)/2 - g2 sum n_i n_{i+1})`` (extgutzwiller.jl:31-35, the call operator; the
    reference's val_and_grad for this ansatz references an undefined name and cannot run -- the gradient here
    is the analytic one of the call operator)."""
    KIND = _capi.ANSATZ_EXT_GUTZWILLER

    @staticmethod
    def _num_parameters(M):
        return 2


class MultinomialAnsatz(GenericAnsatz):
    """``MultinomialAnsatz(H; normalize=false)``: ``(K prod_i 1/n_i!)^p`` (multinomial.jl:17-59)."""
    KIND = _capi.ANSATZ_MULTINOMIAL

    def __init__(self, hamiltonian, normalize: bool = False):
        super().__init__(hamiltonian, normalize=normalize)

    @staticmethod
    def _num_parameters(M):
        return 1


class DensityProfileAnsatz(GenericAnsatz):
    """``exp(-sum_i p_i n_i)`` (densityprofile.jl:8-28); M parameters."""
    KIND = _capi.ANSATZ_DENSITY_PROFILE

    @staticmethod
    def _num_parameters(M):
        return M


class RelativeJastrowAnsatz(GenericAnsatz):
    """``exp(-sum_d p_d sum_k n_k n_{k+d})`` (jastrow.jl:71-103); cld(M, 2) parameters."""
    KIND = _capi.ANSATZ_RELATIVE_JASTROW

    @staticmethod
    def _num_parameters(M):
        return (M + 1) // 2


class JastrowAnsatz(GenericAnsatz):
    """``exp(-sum_{k<=l} p_kl n_k n_l)`` (jastrow.jl:13-58); M (M + 1) / 2 parameters."""
    KIND = _capi.ANSATZ_JASTROW

    @staticmethod
    def _num_parameters(M):
        return M * (M + 1) // 2


def _kind_of(ansatz):
    if isinstance(ansatz, GenericAnsatz) and not isinstance(ansatz, CombinationAnsatz):
        return ansatz.KIND, ansatz.normalize
    if type(ansatz) is GutzwillerAnsatz:
        return _capi.ANSATZ_GUTZWILLER, False
    raise TypeError(f"{type(ansatz).__name__} cannot be a component of a CombinationAnsatz")


class CombinationAnsatz(GenericAnsatz):
    """``ansatz1 + ansatz2`` (combination.jl:1-51): ``alpha * ansatz1 + (1 - alpha) * ansatz2`` with the parameters
    ``(params1..., params2..., alpha)``."""
    KIND = _capi.ANSATZ_COMBINATION

    def __init__(self, ansatz1, ansatz2):
        h1, h2 = ansatz1.hamiltonian, ansatz2.hamiltonian
        if h1 is not h2 and (h1.u, h1.t, getattr(h1, "v", 0.0), h1.address) != (h2.u, h2.t, getattr(h2, "v", 0.0), h2.address):
            raise TypeError("both components must be built on the same Hamiltonian")
        self.ansatz1, self.ansatz2 = ansatz1, ansatz2
        self._kinds = (_kind_of(ansatz1), _kind_of(ansatz2))
        self.hamiltonian = h1
        self.normalize = False
        self.N_PARAMS = ansatz1.N_PARAMS + ansatz2.N_PARAMS + 1  # combination.jl:14
        self._handles = {}

    def handle(self, ctx: Context | None = None) -> _AnsatzHandle:
        ctx = ctx or default_context()
        key = id(ctx)
        if key not in self._handles:
            (k1, n1), (k2, n2) = self._kinds
            self._handles[key] = _AnsatzHandle(self.hamiltonian.model(ctx), (k1, k2), (n1, n2))
        return self._handles[key]

    def __repr__(self):
        return f"CombinationAnsatz({self.ansatz1!r}, {self.ansatz2!r})"


def _add(a1, a2):
    if not isinstance(a2, AbstractAnsatz):
        return NotImplemented
    return CombinationAnsatz(a1, a2)


AbstractAnsatz.__add__ = _add  # Base.:+(a1::AbstractAnsatz, a2::AbstractAnsatz) (combination.jl:13-15)


class GenRunResult:
    """Result of one generic sampling run; attribute names follow ``gwz_vqmc_result``."""

    def __init__(self, r: _capi.GenResult, grad, grad_err, pooled_grad):
        self.val, self.err, self.pooled_val, self.pooled_var = r.val, r.err, r.pooled_val, r.pooled_var
        self.total_time, self.walkers, self.steps, self.hops = r.total_time, r.walkers, r.steps, r.hops
        self.kernel_ms = r.kernel_ms
        self.grad, self.grad_err, self.pooled_grad = grad, grad_err, pooled_grad


class GenWalkers:
    """gwz_gwalkers: n walkers of a generic ansatz (global ids offset..offset+n-1) resident on one device."""

    def __init__(self, ansatz_handle: _AnsatzHandle, n_walkers: int, start_words, seed: int, global_offset: int = 0):
        self.ansatz_handle = ansatz_handle
        self.model = ansatz_handle.model
        self.P = ansatz_handle.P
        self.n, self.offset = int(n_walkers), int(global_offset)
        self._h = C.c_void_p()
        sw = np.ascontiguousarray(start_words, dtype=np.uint64)
        check(lib().gwz_gwalkers_create(ansatz_handle.handle, self.n, self.offset, sw.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_gwalkers_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def steps_done(self) -> int:
        n, s = C.c_uint64(), C.c_uint64()
        check(lib().gwz_gwalkers_count(self._h, C.byref(n), C.byref(s)))
        return s.value

    def addresses(self) -> np.ndarray:
        out = np.zeros((self.n, self.model.words), dtype=np.uint64)
        check(lib().gwz_gwalkers_get_addresses(self._h, _vp(out)))
        return out

    def set_addresses(self, addrs) -> None:
        a = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(self.n, self.model.words)
        check(lib().gwz_gwalkers_set_addresses(self._h, _vp(a)))

    def set_step_counter(self, steps_done: int) -> None:
        check(lib().gwz_gwalkers_set_step_counter(self._h, int(steps_done)))

    def run(self, params, steps: int, burn_in: int = 0, kernel=None, keep_acc: bool = False, pooled: bool = False):
        p = as_params(params, self.P)
        r = _capi.GenResult()
        g, ge, pg = np.zeros(self.P), np.zeros(self.P), np.zeros(self.P)
        flags = (_capi.RUN_KEEP_ACC if keep_acc else 0) | (_capi.RUN_POOLED if pooled else 0)
        check(lib().gwz_gen_vqmc_run(self._h, p.ctypes.data_as(_DP), int(steps), int(burn_in), flags, C.byref(r),
                                     _vp(g), _vp(ge), _vp(pg)))
        return GenRunResult(r, g, ge, pg)

    def run_record(self, params, steps: int, kernel=None, keep_acc: bool = False, want_addresses: bool = True):
        p = as_params(params, self.P)
        steps = int(steps)
        tau, eloc = np.zeros((steps, self.n)), np.zeros((steps, self.n))
        addr = np.zeros((steps, self.n, self.model.words), dtype=np.uint64) if want_addresses else None
        r = _capi.GenResult()
        g, ge, pg = np.zeros(self.P), np.zeros(self.P), np.zeros(self.P)
        check(lib().gwz_gen_vqmc_run_record(self._h, p.ctypes.data_as(_DP), steps, _capi.RUN_KEEP_ACC if keep_acc else 0,
                                            _vp(tau), _vp(eloc), _vp(addr), C.byref(r), _vp(g), _vp(ge), _vp(pg)))
        return GenRunResult(r, g, ge, pg), tau, eloc, addr

    def estimates(self):
        v, g = np.zeros(self.n), np.zeros((self.n, self.P))
        check(lib().gwz_gwalkers_estimates(self._h, _vp(v), _vp(g)))
        return v, g


class GenSweep:
    """gwz_gsweep: LocalEnergyEvaluator(H, ansatz; basis) (localenergy.jl:81-173) on the generic engine."""

    def __init__(self, ansatz_handle: _AnsatzHandle, basis, begin: int, end: int):
        self.ansatz_handle = ansatz_handle
        self.P = ansatz_handle.P
        self.words = ansatz_handle.model.words
        self._h = C.c_void_p()
        if basis is not None:
            check(lib().gwz_gsweep_create(ansatz_handle.handle, _vp(basis), basis.shape[0], 0, 0, C.byref(self._h)))
        else:
            check(lib().gwz_gsweep_create(ansatz_handle.handle, None, 0, int(begin), int(end), C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_gsweep_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def dim(self) -> int:
        d = C.c_uint64()
        check(lib().gwz_gsweep_dim(self._h, C.byref(d)))
        return d.value

    def val(self, params) -> float:
        p, v = as_params(params, self.P), C.c_double()
        check(lib().gwz_gsweep_val_and_grad(self._h, p.ctypes.data_as(_DP), C.byref(v), None))
        return v.value

    def val_and_grad(self, params):
        p, v, g = as_params(params, self.P), C.c_double(), np.zeros(self.P)
        check(lib().gwz_gsweep_val_and_grad(self._h, p.ctypes.data_as(_DP), C.byref(v), _vp(g)))
        return v.value, g

    def terms(self, params, first: int, count: int) -> np.ndarray:
        p, out = as_params(params, self.P), np.zeros((count, 2 + 2 * self.P))
        check(lib().gwz_gsweep_terms(self._h, p.ctypes.data_as(_DP), int(first), int(count), _vp(out)))
        return out

    def states(self, first: int, count: int) -> np.ndarray:
        out = np.zeros((count, self.words), dtype=np.uint64)
        check(lib().gwz_gsweep_states(self._h, int(first), int(count), _vp(out)))
        return out

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(lib().gwz_gsweep_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


class _VectorHandle:
    """gwz_vansatz: the (address, amplitude) pairs of a VectorAnsatz as an HBM hash table on one model."""

    def __init__(self, model, addrs: np.ndarray, vals: np.ndarray):
        self.model = model
        self._h = C.c_void_p()
        check(lib().gwz_vector_ansatz_create(model.handle, _vp(addrs), _vp(vals), addrs.shape[0], C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_vector_ansatz_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def lookup(self, addrs) -> np.ndarray:
        a = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(-1, self.model.words)
        out = np.zeros(a.shape[0])
        check(lib().gwz_vector_ansatz_lookup(self._h, _vp(a), a.shape[0], _vp(out)))
        return out

    def rayleigh(self) -> float:
        v = C.c_double()
        check(lib().gwz_vector_ansatz_rayleigh(self._h, C.byref(v)))
        return v.value


class VectorAnsatz(AbstractAnsatz):
    """``VectorAnsatz(vector)`` (vector.jl:1-12): use a stored vector as a 0-parameter ansatz.  ``vector`` is a mapping
    ``{BoseFS: amplitude}`` (a Rimu DVec) or a pair ``(addresses uint64[n, W], values float64[n])``; addresses it does
    not hold have amplitude 0.  ``kinetic_vqmc(H, VectorAnsatz(v), [])`` samples |v|^2 (test/qmc.jl:22-23)."""
    N_PARAMS = 0
    vector_ansatz = True

    def __init__(self, vector):
        if isinstance(vector, dict):
            keys = list(vector.keys())
            if not keys or not all(isinstance(k, BoseFS) for k in keys):
                raise TypeError("VectorAnsatz needs a non-empty {BoseFS: value} mapping")
            self.addresses = np.stack([k.words for k in keys]).astype(np.uint64)
            self.values = np.array([float(vector[k]) for k in keys])
        else:
            addrs, vals = vector
            self.values = np.ascontiguousarray(vals, dtype=np.float64).ravel()
            self.addresses = np.ascontiguousarray(addrs, dtype=np.uint64).reshape(self.values.size, -1)
        if self.values.size == 0:
            raise ValueError("empty vector")
        self.hamiltonian = None  # bound by the call that uses it (the reference's VectorAnsatz holds no Hamiltonian)
        self._handles = {}

    def handle(self, hamiltonian, ctx: Context | None = None) -> _VectorHandle:
        ctx = ctx or default_context()
        key = (id(hamiltonian), id(ctx))
        if key not in self._handles:
            model = hamiltonian.model(ctx)
            if self.addresses.shape[1] != model.words:
                raise ValueError("the vector's addresses do not match the Hamiltonian's address type")
            self._handles[key] = _VectorHandle(model, self.addresses, self.values)
        return self._handles[key]

    def build_basis(self) -> np.ndarray:
        """``build_basis(va) = collect(keys(va.vector))`` (vector.jl:10)"""
        return self.addresses.copy()

    def __call__(self, addr: BoseFS, params=()) -> float:
        hit = np.all(self.addresses == addr.words[None, :], axis=1)
        return float(self.values[hit][0]) if hit.any() else 0.0

    def val_and_grad(self, addr: BoseFS, params=()):
        return self(addr, params), np.zeros(0)  # abstract.jl:50-52

    def __repr__(self):
        return f"VectorAnsatz({self.values.size} entries)"


class VecWalkers:
    """gwz_vwalkers: walkers of a VectorAnsatz (P = 0)."""

    def __init__(self, vhandle: _VectorHandle, n_walkers: int, start_words, seed: int, global_offset: int = 0):
        self.vhandle, self.model, self.P = vhandle, vhandle.model, 0
        self.n, self.offset = int(n_walkers), int(global_offset)
        self._h = C.c_void_p()
        sw = np.ascontiguousarray(start_words, dtype=np.uint64)
        check(lib().gwz_vwalkers_create(vhandle.handle, self.n, self.offset, sw.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, C.byref(self._h)))
        self._finalizer = weakref.finalize(self, lib().gwz_vwalkers_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def addresses(self) -> np.ndarray:
        out = np.zeros((self.n, self.model.words), dtype=np.uint64)
        check(lib().gwz_vwalkers_get_addresses(self._h, _vp(out)))
        return out

    def run(self, params, steps: int, burn_in: int = 0, kernel=None, keep_acc: bool = False, pooled: bool = False):
        as_params(params if np.size(params) else np.zeros(0), 0)
        r = _capi.GenResult()
        flags = (_capi.RUN_KEEP_ACC if keep_acc else 0) | (_capi.RUN_POOLED if pooled else 0)
        check(lib().gwz_vwalkers_run(self._h, int(steps), int(burn_in), flags, C.byref(r)))
        return GenRunResult(r, np.zeros(0), np.zeros(0), np.zeros(0))

    def run_record(self, params, steps: int, kernel=None, keep_acc: bool = False, want_addresses: bool = True):
        steps = int(steps)
        tau, eloc = np.zeros((steps, self.n)), np.zeros((steps, self.n))
        addr = np.zeros((steps, self.n, self.model.words), dtype=np.uint64) if want_addresses else None
        r = _capi.GenResult()
        check(lib().gwz_vwalkers_run_record(self._h, steps, _capi.RUN_KEEP_ACC if keep_acc else 0, _vp(tau), _vp(eloc),
                                            _vp(addr), C.byref(r)))
        return GenRunResult(r, np.zeros(0), np.zeros(0), np.zeros(0)), tau, eloc, addr


def collect_to_vec(ansatz, params, basis=None, ctx: Context | None = None):
    """``DVec(ansatz, params; basis=build_basis(ansatz))`` (abstract.jl:24-41): the ansatz evaluated on a basis.
    Returns ``(addresses uint64[dim, W], values float64[dim])``; without ``basis`` all C(N+M-1, N) Fock states are
    enumerated on the device (needs N+M-1 <= 64)."""
    if getattr(ansatz, "vector_ansatz", False):
        return ansatz.addresses.copy(), ansatz.values.copy()
    H = ansatz.hamiltonian
    model = H.model(ctx)
    if basis is None:
        from .localenergy import LocalEnergyEvaluator
        basis = LocalEnergyEvaluator(H, ansatz, ctx=ctx).states()
    elif len(basis) and isinstance(basis[0], BoseFS):
        basis = np.stack([b.words for b in basis])
    basis = np.ascontiguousarray(basis, dtype=np.uint64).reshape(-1, model.words)
    if getattr(ansatz, "generic", False):
        vals, _ = ansatz.handle(ctx).eval(basis, params, want_grad=False)
    else:
        vals, _ = model.ansatz_val_and_grad(basis, as_params(params, 1))
    return basis, vals
