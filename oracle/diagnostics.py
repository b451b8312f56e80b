"""TEST INFRASTRUCTURE ONLY — numpy restatement of the chain diagnostics of openmcmc_b200/csrc/diag.cu.

The reference (sede-open/openMCMC) has no ESS / R-hat code at all, so there is nothing to pin this
against: PARITY UNPINNED.  The estimators are the textbook ones — autocorrelation ESS with Geyer's initial monotone
positive sequence (Geyer 1992; Stan reference manual "Effective sample size") and split-R-hat (BDA3 §11.4) — written
here in plain loops so that the CUDA kernels can be checked element for element; `tests/test_oracle_diagnostics.py`
additionally checks them against closed forms (AR(1) processes, i.i.d. draws).
"""

import numpy as np

MAXLAG = 127


def chain_stats(samples, elem_stride=1, n_sel=None, max_lag=MAXLAG):
    """samples [n_iter, n_chains, size] -> [n_chains, n_sel, 8] = n, mean, var, ess, m1, v1, m2, v2."""
    N, C, size = samples.shape
    if n_sel is None:
        n_sel = (size + elem_stride - 1) // elem_stride
    out = np.zeros((C, n_sel, 8))
    L = min(min(max_lag, MAXLAG), N - 1)
    nh = N // 2
    for c in range(C):
        for j in range(n_sel):
            x = samples[:, c, j * elem_stride]
            mean = x.sum() / N
            d = x - mean
            acov = np.array([np.dot(d[: N - l], d[l:]) for l in range(L + 1)])
            c0 = acov[0]
            ess = float(N)
            if c0 > 0 and N > 3:
                s, prev = 0.0, np.inf
                k = 0
                while 2 * k + 1 <= L:
                    pair = (acov[2 * k] + acov[2 * k + 1]) / c0
                    if not pair > 0:
                        break
                    pair = min(pair, prev)
                    s += pair
                    prev = pair
                    k += 1
                tau = max(2 * s - 1, 1.0 / np.log10(N + 9.0))
                ess = N / tau
            h1, h2 = x[:nh], x[N - nh:]
            out[c, j] = [N, mean, c0 / (N - 1) if N > 1 else 0.0, ess,
                         h1.mean() if nh else 0.0, h1.var(ddof=1) if nh > 1 else 0.0,
                         h2.mean() if nh else 0.0, h2.var(ddof=1) if nh > 1 else 0.0]
    return out


def rhat_combine(stats):
    """stats [n_chains_total, n_sel, 8] -> [n_sel, 4] = split-R-hat, total ESS, grand mean, var+."""
    C, n_sel, _ = stats.shape
    out = np.zeros((n_sel, 4))
    for j in range(n_sel):
        nh = np.floor(stats[0, j, 0] / 2)
        means = np.concatenate([stats[:, j, 4], stats[:, j, 6]])
        W = np.concatenate([stats[:, j, 5], stats[:, j, 7]]).mean()
        b_over_n = means.var(ddof=1)
        var_plus = (nh - 1) / nh * W + b_over_n
        out[j] = [np.sqrt(var_plus / W) if W > 0 else np.nan, stats[:, j, 3].sum(), stats[:, j, 1].mean(), var_plus]
    return out


def rank_normalize(samples, elem_stride=1, n_sel=None, pooled=False):
    """Vehtari et al. (2021), eq. (14): z = Phi^-1((r - 3/8) / (S + 1/4)), r = average rank among the S draws ranked
    together (one chain's series, or all chains pooled).  samples [n_iter, n_chains, size] -> z [n_iter, n_chains, n_sel]."""
    from scipy import stats

    N, C, size = samples.shape
    if n_sel is None:
        n_sel = (size + elem_stride - 1) // elem_stride
    z = np.empty((N, C, n_sel))
    for j in range(n_sel):
        x = samples[:, :, j * elem_stride]
        if pooled:
            r = stats.rankdata(x.ravel(), method="average").reshape(N, C)
            z[:, :, j] = stats.norm.ppf((r - 0.375) / (N * C + 0.25))
        else:
            for c in range(C):
                r = stats.rankdata(x[:, c], method="average")
                z[:, c, j] = stats.norm.ppf((r - 0.375) / (N + 0.25))
    return z
