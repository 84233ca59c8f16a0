"""Independent checker for the error-bar path (a8 / f1): resample + blocking analysis with the M test.

TEST INFRASTRUCTURE.  Written from the published sources only -- it shares no code with csrc/gwz_stats.cpp,
csrc/gwz_blocking.cu or oracle/gutz_oracle.c, and uses a different formulation (whole-array numpy, two-pass
moments) than the product (streaming, centred one-pass moments).

Sources
  [FP]  H. Flyvbjerg, H. G. Petersen, J. Chem. Phys. 91, 461 (1989): repeated pair averaging; the standard
        error of the mean at blocking level j is sqrt(var_j / n_j), its own uncertainty se / sqrt(2 (n_j - 1)).
  [J]   M. Jonsson, Phys. Rev. E 98, 043304 (2018): automated choice of the level.  Theorem: under H0
        (level k and above uncorrelated) M_k = sum_{j >= k} M_j,
        M_j = n_j [ (n_j - 1) s_j^2 / n_j^2 + gamma_j(1) ]^2 / s_j^4, is chi-square distributed; the paper's
        Python listing (`block(x)`) tests `M[k] < q[k]` with q = the 0.99 chi-square quantiles for 1..30
        degrees of freedom, INDEXED BY THE LEVEL k ITSELF (k counted from the unblocked series), and uses the
        large-n form M_j = n_j (gamma_j / s_j^2)^2.
  Rimu.StatsTools (`blocking_analysis`, `blocks_with_m`, `mtest`; the reference calls it at
        src/kinetic-vqmc/state.jl:53,273) follows [J]: same quantile table, `m[k] < q[k]`, k = 1 .. levels-1,
        finite-n M_j with `var(v; corrected)` and `autocovariance(v, 1; corrected)` (both divided by n-1 for
        the default corrected=true), std_err = sqrt(var/n), std_err_err = std_err / sqrt(2(n-1)).  Rimu's source
        is NOT available in this environment (un-vendored dependency): that description is
        from memory of the package and stays "parity unpinned"; what IS pinned here is the published listing.

Decision documented in DESIGN.md section 2: the quantile is indexed by the level (q[k]), as printed in [J]
and as Rimu does -- not by the number of remaining levels (round 1 used q[levels - k], which is the
statistically motivated reading of the theorem but is not what either source computes).
"""
from __future__ import annotations

import math

import numpy as np

# the "magic numbers" of the listing in [J]: chi-square 0.99 quantiles, 1..30 degrees of freedom
Q99 = np.array([6.634897, 9.210340, 11.344867, 13.276704, 15.086272, 16.811894, 18.475307, 20.090235, 21.665994,
                23.209251, 24.724970, 26.216967, 27.688250, 29.141238, 30.577914, 31.999927, 33.408664, 34.805306,
                36.190869, 37.566235, 38.932173, 40.289360, 41.638398, 42.979820, 44.314105, 45.641683, 46.962942,
                48.278236, 49.587884, 50.892181])


def jonsson_block(x):
    """The listing of [J] (appendix), for len(x) = 2^d.  Returns (mean, variance of the mean, k, M)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    d = int(math.log2(n))
    assert n == 1 << d, "the published listing assumes a power-of-two length"
    s, gamma = np.zeros(d), np.zeros(d)
    mu = np.mean(x)
    for i in range(d):
        n = len(x)
        gamma[i] = np.sum((x[0:n - 1] - mu) * (x[1:n] - mu)) / n
        s[i] = np.var(x)
        x = 0.5 * (x[0::2] + x[1::2])
    M = np.cumsum(((gamma / s) ** 2 * 2.0 ** np.arange(1, d + 1)[::-1])[::-1])[::-1]
    k = d - 1
    for kk in range(d):
        if M[kk] < Q99[kk]:
            k = kk
            break
    return mu, s[k] / 2 ** (d - k), k, M


def fp_levels(x, corrected=True):
    """[FP] levels of an arbitrary-length series (an odd last element is dropped at each pair averaging).
    Per level: n, mean, standard error, its uncertainty, and the finite-n M_j of [J]'s theorem."""
    v = np.asarray(x, dtype=np.float64).copy()
    levels = int(math.floor(math.log2(len(v))))
    out = dict(n=[], mean=[], se=[], see=[], mj=[])
    for _ in range(levels):
        n = len(v)
        mean = v.sum() / n
        dv = v - mean
        ddof = 1 if corrected else 0
        var = np.dot(dv, dv) / (n - ddof)
        gamma = np.dot(dv[:-1], dv[1:]) / (n - ddof)
        se = math.sqrt(var / n)
        out["n"].append(n)
        out["mean"].append(mean)
        out["se"].append(se)
        out["see"].append(se / math.sqrt(2.0 * (n - 1)))
        with np.errstate(divide="ignore", invalid="ignore"):
            out["mj"].append(np.float64(n) * ((n - 1) * var / n ** 2 + gamma) ** 2 / np.float64(var) ** 2)
        v = 0.5 * (v[0:2 * (n // 2):2] + v[1:2 * (n // 2):2])
    return {k: np.array(a) for k, a in out.items()}


def blocking_mtest(x, corrected=True):
    """[FP] levels + the M test of [J] with the listing's indexing.  Returns a dict with
    mean, err, err_err, k (1-based level, -1 = "more data needed"), blocks, and the level table."""
    x = np.asarray(x, dtype=np.float64)
    if len(x) == 1:
        return dict(mean=float(x[0]), err=math.nan, err_err=math.nan, k=0, blocks=1, levels=None)
    lv = fp_levels(x, corrected)
    mj = lv["mj"]
    M = np.cumsum(mj[::-1])[::-1]          # M_k = sum_{j >= k} M_j
    k = -1
    for kk in range(len(M) - 1):            # the top level is never accepted
        if M[kk] < Q99[kk]:                 # NaN (constant series) compares false
            k = kk + 1
            break
    res = dict(mean=float(lv["mean"][0]), err=math.nan, err_err=math.nan, k=k, blocks=0, levels=lv, M=M)
    if k > 0:
        res.update(err=float(lv["se"][k - 1]), err_err=float(lv["see"][k - 1]), blocks=int(lv["n"][k - 1]))
    return res


def resample_integral(times, values, length=None):
    """Equal-time resampling (src/kinetic-vqmc/state.jl:183-231) from its DEFINITION: the series is the step
    function that holds values[k] for times[k]; rescale the time axis to [0, length] and average over unit slots.
    Output slot i = F(i) - F(i-1) with F the running integral (piecewise linear, np.interp)."""
    t = np.asarray(times, dtype=np.float64)
    v = np.asarray(values, dtype=np.float64)
    n_out = len(t) if length is None else int(length)
    edges = np.concatenate([[0.0], np.cumsum(t)]) * (n_out / np.sum(t))
    F = np.concatenate([[0.0], np.cumsum(v * np.diff(edges))])
    grid = np.arange(n_out + 1, dtype=np.float64)
    grid[-1] = edges[-1]
    return np.diff(np.interp(grid, edges, F))


def ar1(rng, n, phi, shape=()):
    """AR(1) series x_t = phi x_{t-1} + eps_t, started in the stationary distribution; unit innovations."""
    x = np.empty(shape + (n,))
    x[..., 0] = rng.standard_normal(shape) / math.sqrt(1.0 - phi * phi)
    eps = rng.standard_normal(shape + (n,))
    for i in range(1, n):
        x[..., i] = phi * x[..., i - 1] + eps[..., i]
    return x


def ar1_error_of_mean(n, phi):
    """exact standard error of the mean of n consecutive samples of the stationary AR(1) process above"""
    var_x = 1.0 / (1.0 - phi * phi)
    # Var(mean) = var_x / n * [1 + 2 sum_{h=1}^{n-1} (1 - h/n) phi^h]
    h = np.arange(1, n)
    return math.sqrt(var_x / n * (1.0 + 2.0 * np.sum((1.0 - h / n) * phi ** h)))
