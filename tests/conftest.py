import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)  # sibling helper modules (sparse_models)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices (skipped otherwise)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build()
    return om


@pytest.fixture(scope="session")
def gw():
    """The product package; on a GPU box the CUDA library must be the thing that runs."""
    import gutzwiller_b200 as g
    return g


def random_onr(rng, N, M, concentration=1.0):
    """Random occupation vector; small concentration -> clumped states (large occupations)."""
    p = rng.dirichlet(np.full(M, concentration))
    return rng.multinomial(N, p)


def edge_onrs(N, M):
    out = []
    for pos in (0, M - 1, M // 2):
        o = np.zeros(M, dtype=int); o[pos] = N; out.append(o)             # all bosons on one site
    o = np.zeros(M, dtype=int); o[0] = N // 2; o[M - 1] = N - N // 2; out.append(o)  # across the periodic bond
    o = np.zeros(M, dtype=int); o[::2] = 1; o[0] += N - o.sum(); out.append(o)      # alternating + pile
    ff, ex = divmod(N, M)
    out.append(np.array([ff + (1 if i < ex else 0) for i in range(M)]))  # near uniform
    o = np.zeros(M, dtype=int); o[1] = N - 1; o[2 % M] += 1; out.append(o)
    if M >= 6 and N >= 9:
        o = np.zeros(M, dtype=int); o[M - 2] = 4; o[M - 1] = 0; o[0] = 5; o[1] = N - 9; out.append(o)
    return [x for x in out if x.sum() == N and (x >= 0).all()]
