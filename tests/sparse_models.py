"""Test operators for the stored-operator path (TEST INFRASTRUCTURE): small Hamiltonians materialised as
(row_ptr, col, melem, diag) with rows in a fixed `offdiagonals` order, the way the Julia shim would materialise
`build_basis(H)` + `offdiagonals(H, k)` for any Rimu Hamiltonian (test/qmc.jl:10-14 uses HubbardReal1D,
HubbardMom1D and a two-component HubbardRealSpace).

Rimu's operator code is not available here, so the momentum-space and two-component operators below are written from
the textbook second-quantised Hamiltonians; the hop ORDER inside a row is this file's own (documented per builder).
They are checked for Hermiticity and, for the momentum-space chain, against the real-space spectrum
(tests/test_oracle_sparse.py)."""
from __future__ import annotations

import itertools
import math

import numpy as np


def _csr(rows, diag):
    row_ptr = np.zeros(len(rows) + 1, dtype=np.uint64)
    row_ptr[1:] = np.cumsum([len(r) for r in rows])
    col = np.array([c for r in rows for c, _ in r], dtype=np.uint32)
    mel = np.array([m for r in rows for _, m in r], dtype=np.float64)
    return row_ptr, col, mel, np.asarray(diag, dtype=np.float64)


def hubbard_real1d(oracle):
    """HubbardReal1D through part 1 of the oracle: BFS basis (App. A.5), rows in offdiagonals order (App. A.4).
    Returns (row_ptr, col, melem, diag, basis_words)."""
    basis = oracle.build_basis()
    index = {tuple(int(x) for x in b): i for i, b in enumerate(basis)}
    rows, diag = [], []
    for b in basis:
        diag.append(oracle.diagonal_element(b))
        rows.append([(index[tuple(int(x) for x in a)], m) for a, m in oracle.offdiagonals(b)])
    return (*_csr(rows, diag), basis)


def _closure(start, neighbours):
    """breadth-first closure (build_basis): returns (states, index, rows)"""
    states, index, rows = [start], {start: 0}, []
    q = 0
    while q < len(states):
        row = []
        for target, m in neighbours(states[q]):
            if target not in index:
                index[target] = len(states)
                states.append(target)
            row.append((index[target], m))
        rows.append(row)
        q += 1
    return states, rows


def hubbard_mom1d(onr, u=1.0, t=1.0):
    """Bosonic Hubbard chain in momentum space (the operator behind Rimu's HubbardMom1D):
        H = sum_k eps_k n_k + u/(2M) sum_{p,q,r,s} a+_r a+_s a_q a_p delta(r+s, p+q),  eps_k = -2t cos(2 pi k / M).
    Row order: ordered pairs (p, q) ascending, momentum transfer d = 1..M-1 (p -> p+d, q -> q-d); the two orderings
    of a pair are separate entries (duplicate targets), like an operator that lists every excitation."""
    onr = tuple(int(x) for x in onr)
    M = len(onr)
    eps = [-2.0 * t * math.cos(2.0 * math.pi * k / M) for k in range(M)]

    def neighbours(n):
        out = []
        for p in range(M):
            for q in range(M):
                if n[p] == 0 or n[q] == 0 or (p == q and n[p] < 2):
                    continue
                for d in range(1, M):
                    r, s = (p + d) % M, (q - d) % M
                    m = list(n)
                    amp = math.sqrt(m[p])
                    m[p] -= 1
                    amp *= math.sqrt(m[q])
                    m[q] -= 1
                    m[s] += 1
                    amp *= math.sqrt(m[s])
                    m[r] += 1
                    amp *= math.sqrt(m[r])
                    m = tuple(m)
                    if m != n:
                        out.append((m, u / (2.0 * M) * amp))
        return out

    def diagonal(n):
        N = sum(n)
        kin = sum(e * x for e, x in zip(eps, n))
        # terms of the interaction that return to n: (r,s) = (p,q) and (r,s) = (q,p)
        same = sum(x * (x - 1) for x in n)
        cross = sum(n[p] * n[q] for p in range(M) for q in range(M) if p != q)
        assert N >= 0
        return kin + u / (2.0 * M) * (same + 2.0 * cross)

    states, rows = _closure(onr, neighbours)
    return (*_csr(rows, [diagonal(s) for s in states]), states)


def hubbard_2c_lattice(up, down, dims=(2, 2), u=1.0, t=1.0):
    """Two-component fermions on a periodic lattice (the operator behind
    HubbardRealSpace(FermiFS2C(up, down); geometry=PeriodicBoundaries(dims))): nearest-neighbour hopping -t with
    Jordan-Wigner signs inside each component, on-site u n_up n_down.  Row order: component (up, down), source site
    ascending, direction (+x, -x, +y, -y); on a length-2 axis +d and -d reach the same site and both are listed."""
    up, down = tuple(up), tuple(down)
    nsite = len(up)
    assert int(np.prod(dims)) == nsite
    coords = list(itertools.product(*[range(d) for d in dims]))  # last axis fastest
    site_of = {c: i for i, c in enumerate(coords)}

    def hop_sign(occ, i, j):
        lo, hi = (i, j) if i < j else (j, i)
        return -1.0 if sum(occ[lo + 1:hi]) % 2 else 1.0

    def neighbours(state):
        out = []
        for comp in (0, 1):
            occ = state[comp]
            for i in range(nsite):
                if not occ[i]:
                    continue
                for axis in range(len(dims)):
                    for step in (+1, -1):
                        c = list(coords[i])
                        c[axis] = (c[axis] + step) % dims[axis]
                        j = site_of[tuple(c)]
                        if j == i or occ[j]:
                            continue
                        new = list(occ)
                        new[i], new[j] = 0, 1
                        tgt = (tuple(new), state[1]) if comp == 0 else (state[0], tuple(new))
                        out.append((tgt, -t * hop_sign(occ, i, j)))
        return out

    states, rows = _closure((up, down), neighbours)
    diag = [u * sum(a * b for a, b in zip(s[0], s[1])) for s in states]
    return (*_csr(rows, diag), states)


def random_operator(dim, mean_row, seed=0, max_row=None, signed=True):
    """A random symmetric-pattern operator with ragged rows (some much longer than the mean, duplicates included)."""
    rng = np.random.default_rng(seed)
    max_row = max_row or 6 * mean_row
    sets = [[] for _ in range(dim)]
    for k in range(dim):
        deg = int(min(max_row, max(1, rng.geometric(1.0 / mean_row))))
        for l in rng.integers(0, dim, size=deg):
            if l == k:
                continue
            m = rng.normal() if signed else -abs(rng.normal())
            sets[k].append((int(l), m))
            sets[int(l)].append((k, m))
    for k in range(dim):  # no empty rows: the walk must always be able to leave
        if not sets[k]:
            l = (k + 1) % dim
            sets[k].append((l, -1.0))
            sets[l].append((k, -1.0))
    return _csr(sets, rng.normal(size=dim))


def dense(row_ptr, col, melem, diag):
    dim = len(diag)
    H = np.diag(np.asarray(diag, dtype=np.float64))
    rp = row_ptr.astype(np.int64)
    for k in range(dim):
        for i in range(rp[k], rp[k + 1]):
            H[k, col[i]] += melem[i]
    return H
