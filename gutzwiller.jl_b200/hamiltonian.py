"""Host-side mirrors of the Rimu.jl types the reference API is called with.

Only what the hot path needs: ``BoseFS`` addresses, ``near_uniform``, ``HubbardReal1D`` and the
``GutzwillerAnsatz`` (src/ansatz/gutzwiller.jl:14-36) / ``AbstractAnsatz`` traits
(src/ansatz/abstract.jl:15-22).  These objects hold parameters; all arithmetic happens on the GPU
through :class:`runtime.Model`.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .runtime import Context, Model, default_context


def _words_for(bits: int) -> int:
    return (bits + 63) // 64


class BoseFS:
    """``BoseFS{N,M}``: N bosons in M modes as a unary bitstring of N+M-1 bits.

    ``BoseFS((1,1,1,1))`` / ``BoseFS(1,1,1,1)`` build from occupation numbers; ``BoseFS.from_words``
    wraps raw 64-bit words (word 0 = least significant).
    """

    __slots__ = ("N", "M", "words")

    def __init__(self, *onr):
        if len(onr) == 1 and not np.isscalar(onr[0]):
            onr = tuple(onr[0])
        onr = [int(x) for x in onr]
        if len(onr) < 2 or any(x < 0 for x in onr):
            raise ValueError("BoseFS needs >= 2 modes with non-negative occupations")
        self.N, self.M = sum(onr), len(onr)
        if self.N < 1:
            raise ValueError("BoseFS needs at least one particle")
        bits = self.N + self.M - 1
        if bits > 64 * _capi.MAX_WORDS - 5:
            raise ValueError("N+M-1 must be <= 251")
        val, pos = 0, 0
        for n in onr:
            val |= ((1 << n) - 1) << pos
            pos += n + 1
        W = _words_for(bits)
        self.words = np.array([(val >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(W)], dtype=np.uint64)

    @classmethod
    def from_words(cls, N: int, M: int, words) -> "BoseFS":
        self = object.__new__(cls)
        self.N, self.M = int(N), int(M)
        W = _words_for(N + M - 1)
        w = np.asarray(words, dtype=np.uint64).ravel()
        if w.size < W:
            raise ValueError("not enough words for this BoseFS{N,M}")
        self.words = w[:W].copy()
        if bin(self._int()).count("1") != N or self._int() >> (N + M - 1):
            raise ValueError("words do not encode N particles in N+M-1 bits")
        return self

    def _int(self) -> int:
        return sum(int(w) << (64 * i) for i, w in enumerate(self.words))

    def onr(self) -> tuple:
        v, out, run = self._int(), [], 0
        for p in range(self.N + self.M - 1):
            if (v >> p) & 1:
                run += 1
            else:
                out.append(run)
                run = 0
        out.append(run)
        return tuple(out)

    @property
    def num_particles(self) -> int:
        return self.N

    @property
    def num_modes(self) -> int:
        return self.M

    def __eq__(self, other):
        return isinstance(other, BoseFS) and (self.N, self.M) == (other.N, other.M) and \
            np.array_equal(self.words, other.words)

    def __hash__(self):
        return hash((self.N, self.M, self._int()))

    def __repr__(self):
        return 'fs"|' + " ".join(str(n) for n in self.onr()) + '⟩"'


def near_uniform(N: int, M: int) -> BoseFS:
    """``near_uniform(BoseFS{N,M})``: filling factor everywhere, the first N % M modes get one more."""
    ff, extras = divmod(int(N), int(M))
    return BoseFS([ff + (1 if i < extras else 0) for i in range(M)])


def num_modes(addr: BoseFS) -> int:
    return addr.M


class HubbardReal1D:
    """``HubbardReal1D(addr; u=1.0, t=1.0)``: periodic 1-D Bose-Hubbard chain (App. A.2)."""

    def __init__(self, addr: BoseFS, u: float = 1.0, t: float = 1.0):
        if not isinstance(addr, BoseFS):
            raise TypeError("HubbardReal1D needs a BoseFS starting address")
        self.address, self.u, self.t = addr, float(u), float(t)
        self._models: dict[int, Model] = {}

    def model(self, ctx: Context | None = None) -> Model:
        ctx = ctx or default_context()
        key = id(ctx)
        if key not in self._models:
            self._models[key] = Model(self.address.N, self.address.M, self.u, self.t, ctx, v=getattr(self, "v", 0.0))
        return self._models[key]

    def __repr__(self):
        return f"HubbardReal1D({self.address!r}; u={self.u}, t={self.t})"


def starting_address(H: HubbardReal1D) -> BoseFS:
    return H.address


class AbstractAnsatz:
    """``AbstractAnsatz{K,V,N}`` contract (src/ansatz/abstract.jl:15-22)."""
    N_PARAMS = 0


class GutzwillerAnsatz(AbstractAnsatz):
    """``G(|f>; g) = exp(-g <f|H|f>)`` (src/ansatz/gutzwiller.jl:14-36); one parameter g."""
    N_PARAMS = 1

    def __new__(cls, hamiltonian=None):
        # GutzwillerAnsatz(ExtendedHubbardReal1D) runs on the ansatz-generic engine (ansatz.py)
        if cls is GutzwillerAnsatz and getattr(hamiltonian, "v", 0.0):
            from .ansatz import GenericGutzwillerAnsatz
            return GenericGutzwillerAnsatz(hamiltonian)
        return object.__new__(cls)

    def __init__(self, hamiltonian: HubbardReal1D):
        if not isinstance(hamiltonian, HubbardReal1D):
            raise TypeError("GutzwillerAnsatz needs a HubbardReal1D / ExtendedHubbardReal1D Hamiltonian")
        self.hamiltonian = hamiltonian

    def __call__(self, addr: BoseFS, params) -> float:
        return self.val_and_grad(addr, params)[0]

    def val_and_grad(self, addr: BoseFS, params):
        p = as_params(params, 1)
        val, grad = self.hamiltonian.model().ansatz_val_and_grad(addr.words[None, :], p)
        return float(val[0]), grad[0].copy()

    def __repr__(self):
        return f"GutzwillerAnsatz({self.hamiltonian!r})"


def num_parameters(ansatz: AbstractAnsatz) -> int:
    return ansatz.N_PARAMS


def keytype(ansatz: AbstractAnsatz):
    return type(ansatz.hamiltonian.address)


def valtype(ansatz: AbstractAnsatz):
    return float


def as_params(params, n: int) -> np.ndarray:
    """Accept a scalar, list, tuple or array like ``SVector{N,T}(params)`` does (state.jl:25)."""
    p = np.atleast_1d(np.asarray(params, dtype=np.float64)).ravel() if np.size(params) else np.zeros(0)
    if p.size != n:
        raise ValueError(f"expected {n} parameter(s), got {p.size}")
    return np.ascontiguousarray(p)
