"""CPU: host-side logic of the device engine that needs no GPU — structure classification of precision matrices
(engine.classify_matrix), the chain-block policy of MCMC.run_mcmc (upload_blocks), state coercion (mcmc.py:63-76), the
lazy final-state entry, and the element order of the kernels' division-free 2-D loops (restated in Python)."""

import numpy as np
import pytest
from scipy import sparse


def test_classify_matrix_structures():
    from openmcmc_b200.engine import classify_matrix

    n = 200
    rng = np.random.default_rng(0)
    d = rng.random(n) + 0.5
    e = -rng.random(n - 1)
    assert classify_matrix(sparse.identity(n, format="csc"))[0] == "eye"
    assert classify_matrix(sparse.identity(n, format="coo"))[0] == "eye"
    kind, main, off = classify_matrix(sparse.diags([d], [0], format="csr"))
    assert kind == "diag" and np.array_equal(main, d) and off is None
    kind, main, off = classify_matrix(sparse.diags([e, d, e], [-1, 0, 1], format="csc"))
    assert kind == "tridiag" and np.array_equal(main, d) and np.array_equal(off, e)
    # explicit zeros and duplicate entries are canonicalised before the structure is read
    T = sparse.diags([e, d, e], [-1, 0, 1], format="coo")
    dup = sparse.coo_matrix((np.concatenate([T.data, [0.25, 0.0]]), (np.concatenate([T.row, [3, 0]]),
                                                                     np.concatenate([T.col, [3, n - 1]]))), shape=(n, n))
    kind, main, off = classify_matrix(dup)
    d2 = d.copy()
    d2[3] += 0.25
    assert kind == "tridiag" and np.allclose(main, d2) and np.array_equal(off, e)
    # non-symmetric tridiagonal: refused; wider bands: dense when small, refused when large
    with pytest.raises(NotImplementedError):
        classify_matrix(sparse.diags([e, d, 2 * e], [-1, 0, 1], format="csc"))
    small = sparse.csc_matrix(np.array([[2.0, 1, 0.5], [1, 2, 1], [0.5, 1, 2]]))
    assert classify_matrix(small)[0] == "dense"
    assert classify_matrix(sparse.diags([e[:-1], e, d, e, e[:-1]], [-2, -1, 0, 1, 2], format="csc"))[0] == (
        "dense" if n <= 512 else "banded")
    nb = 600
    with pytest.raises(NotImplementedError):
        classify_matrix(sparse.diags([np.ones(nb - 2), np.ones(nb - 1), 4 * np.ones(nb), np.ones(nb - 1), np.ones(nb - 2)],
                                     [-2, -1, 0, 1, 2], format="csc"))
    # dense inputs
    assert classify_matrix(np.eye(5))[0] == "eye"
    assert classify_matrix(np.diag(d[:5]))[0] == "diag"
    assert classify_matrix(np.array([[2.0, 1], [1, 2]]))[0] == "dense"
    big = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    kind, main, off = classify_matrix(big)
    assert kind == "tridiag" and np.array_equal(main, d) and np.array_equal(off, e)


def _regression_pieces(C, n, p, X):
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    mdl = Model([
        Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
        Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
        Gamma("tau", shape="a_tau", rate="b_tau"),
        Gamma("lambda", shape="a_lambda", rate="b_lambda")])
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
    state = {"y": np.zeros((C, n, 1)), "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"),
             "tau": 1.0, "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01,
             "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}
    return mdl, samplers, state


def test_upload_block_policy_and_state_coercion():
    from openmcmc_b200.mcmc import MCMC

    C, n, p = 12, 5, 2
    mdl, samplers, state = _regression_pieces(C, n, p, np.zeros((C, n, p)))
    M = MCMC(state, samplers, model=mdl, n_chains=C)
    # reference coercion (mcmc.py:65-76): scalars become (1,1) float64 arrays, sparse entries are left alone
    assert M.state["tau"].shape == (1, 1) and M.state["tau"].dtype == np.float64
    assert sparse.issparse(M.state["P_tau"]) and M.state["beta"].shape == (p, 1)
    assert M._n_blocks() == 1                                   # automatic: a few hundred bytes of host input
    assert MCMC(state, samplers, model=mdl, n_chains=C, upload_blocks=4)._n_blocks() == 4
    assert MCMC(state, samplers, model=mdl, n_chains=C, upload_blocks=100)._n_blocks() == 6   # >= 2 chains per block
    assert MCMC(state, samplers, model=mdl, n_chains=C, upload_blocks=4, probes=True)._n_blocks() == 1
    assert MCMC(state, samplers, model=mdl, n_chains=C, upload_blocks=4,
                debug_draws={"beta": {"z": np.zeros((1, C, p))}})._n_blocks() == 1
    assert MCMC(state, samplers, model=mdl, n_chains=1, upload_blocks=4)._n_blocks() == 1
    # automatic policy: one block per 2.7 GB of per-chain host input, at most 16 (a strided view stands in for 21 GB)
    big = np.lib.stride_tricks.as_strided(np.zeros(1), shape=(4096, 10_000, 64), strides=(0, 0, 0))
    st2 = dict(state, X=big, y=np.zeros((4096, 1, 1)))
    auto = MCMC.__new__(MCMC)
    auto.state, auto.samplers, auto.n_chains, auto.upload_blocks, auto.debug_draws, auto.probes = st2, samplers, 4096, None, None, False
    assert auto._n_blocks() == 7
    auto.upload_blocks = 0
    assert auto._n_blocks() == 1


def test_lazy_host_array_fetches_once_and_only_when_read():
    from openmcmc_b200.mcmc import LazyHostArray

    calls = []

    def fetch():
        calls.append(1)
        return np.arange(24.0).reshape(2, 3, 4)

    a = LazyHostArray(fetch, (2, 3, 4))
    assert a.shape == (2, 3, 4) and a.ndim == 3 and len(a) == 2 and not calls
    assert np.asarray(a).sum() == 276.0 and a[1, 2, 3] == 23.0 and len(calls) == 1
    assert np.asarray(a, dtype=np.float32).dtype == np.float32 and len(calls) == 1


@pytest.mark.parametrize("rows,cols,nt", [(16, 33, 128), (16, 129, 128), (7, 5, 128), (33, 33, 128), (1, 1, 128), (64, 64, 128)])
def test_division_free_2d_loop_visits_every_element_once_in_flat_order(rows, cols, nt):
    """rj_for2d / rm_for2d (rj.cu, rj_moves.cu): thread t visits e = t, t + NT, ... of the flattened rows x cols block
    with (i, c) advanced incrementally (one carry at most per step)."""
    seen = np.zeros((rows, cols), dtype=int)
    q, rr = nt // cols, nt - (nt // cols) * cols
    for t in range(nt):
        i, c = t // cols, t - (t // cols) * cols
        e = t
        while i < rows:
            assert i * cols + c == e
            seen[i, c] += 1
            c += rr
            i += q
            if c >= cols:
                c -= cols
                i += 1
            e += nt
    assert np.all(seen == 1)


def test_precision_builders_match_the_rw1_definition():
    """gmrf.precision_irregular / precision_temporal (gmrf.py:351-411): diagonal 1/d_{i-1} + 1/d_i, off-diagonal -1/d_i,
    sparse CSC or dense, a single location gives [[1]]; checked against the oracle's diagonals and the defining
    property x' P x = sum (x_{i+1} - x_i)^2 / d_i."""
    import pandas as pd

    from openmcmc_b200 import gmrf
    from oracle import gmrf as ogmrf

    rng = np.random.default_rng(3)
    s = np.cumsum(rng.exponential(size=40)) + 5.0
    P = gmrf.precision_irregular(s)
    assert sparse.issparse(P) and P.format == "csc" and P.shape == (40, 40)
    d, e = ogmrf.precision_irregular_diagonals(s)
    assert np.array_equal(P.diagonal(0), d) and np.array_equal(P.diagonal(1), e) and np.array_equal(P.diagonal(-1), e)
    Pd = gmrf.precision_irregular(s.reshape(-1, 1), is_sparse=False)
    assert isinstance(Pd, np.ndarray) and np.array_equal(Pd, P.toarray())
    x = rng.standard_normal(40)
    assert np.isclose(x @ (P @ x), np.sum(np.diff(x) ** 2 / np.diff(s)), rtol=1e-12)
    assert np.allclose(P @ np.ones(40), 0.0, atol=1e-12)          # constants are in the null space (rank n - 1)
    assert np.array_equal(gmrf.precision_irregular(np.array([2.5])), np.array([[1]]))
    t = pd.date_range(start="2022-04-01T01:00:00", end="2022-04-01T01:01:00", periods=100)
    Pt = gmrf.precision_temporal(t)
    assert np.allclose(Pt.diagonal(1), -99.0 / 60.0) and np.isclose(Pt[0, 0], 99.0 / 60.0)
    assert np.allclose(gmrf.precision_temporal(t, unit_length=60.0).diagonal(1), -99.0)
