"""CPU: the numpy ReversibleJump oracle (oracle/rj.py) replays the golden steps recorded from the live reference
(tests/golden/rj_*.npz, generator tests/golden/make_golden.py).

Tolerance: the reference solves (S + 1e-10 I) G = S[:, cols] by LU (reversible_jump.py:239-242), whose rounding error
grows like cond(S) * 1e-16; the oracle evaluates the same G as (I - 1e-10 (S + 1e-10 I)^-1)[:, cols].  The goldens record
cond(S) for every step and the comparison allows 1e-9 + 4 * cond * 2.2e-16 on the coefficients (SURVEY §7 "state
tolerance as kappa * eps")."""

import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "rj_*.npz")))


def model_of(g):
    lim = g["match_limits"]
    return dict(X=g["X"], y=g["y"] if str(g["response"]) == "normal" else None, tau_y=float(g["tau_y"]),
                tau_beta=float(g["tau_beta"]), mu_beta=float(g["mu_beta"]), rho=float(g["rho"]),
                a_omega=float(g["a_omega"]) if g["with_omega"] else None, b_omega=float(g["b_omega"]),
                theta_lo=float(g["theta_lo"]), theta_hi=float(g["theta_hi"]), n_max=int(g["n_max"]),
                birth_probability=float(g["birth_probability"]), match_scale=float(g["match_scale"]),
                match_limits=None if np.isnan(lim[0]) else (float(lim[0]), float(lim[1])))


def state_of(g, it, which="before"):
    from oracle import rj

    n = int(g["n_" + which][it])
    th, om, be = (g[k + "_" + which][it][:n] for k in ("theta", "omega", "beta"))
    return dict(n=n, theta=th, omega=om, beta=be, B=rj.make_basis(g["X"], th, om))


def draws_of(g, it):
    return dict(u_move=g["u_move"][it], theta_new=g["theta_new"][it], omega_new=g["omega_new"][it],
                beta_new=g["beta_new"][it], del_index=g["del_index"][it], u_accept=g["u_accept"][it])


def tol_of(g, it):
    return 1e-9 + 4 * g["cond"][it] * 2.2e-16


@pytest.mark.parametrize("name", NAMES)
def test_rj_oracle_replays_reference_steps(name):
    from oracle import rj

    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    m = model_of(g)
    n_steps = g["birth"].size
    births = deaths = accepts = 0
    for it in range(n_steps):
        st = state_of(g, it)
        new, info = rj.rj_step(m, st, draws_of(g, it))
        tol = tol_of(g, it)
        assert info["birth"] == bool(g["birth"][it])
        assert info["accepted"] == bool(g["accepted"][it]), (it, info["log_accept"], g["log_accept"][it])
        for k in ("lq_fwd", "lq_rev", "log_accept"):
            if np.isnan(g[k][it]):
                assert np.isnan(info[k])
            else:
                np.testing.assert_allclose(info[k], g[k][it], rtol=1e-9, atol=50 * tol, err_msg=f"{k} step {it}")
        n_after = int(g["n_after"][it])
        assert new["n"] == n_after
        np.testing.assert_allclose(new["theta"], g["theta_after"][it][:n_after], rtol=1e-12)
        np.testing.assert_allclose(new["beta"], g["beta_after"][it][:n_after], rtol=1e-9, atol=tol)
        births += info["birth"]
        deaths += not info["birth"]
        accepts += info["accepted"]
    assert births > 0 and deaths > 0 and accepts > 0


def test_move_probabilities_edge_cases():
    """reversible_jump.py:361-373"""
    from oracle import rj

    assert rj.move_probabilities(5, 10, 0.3, True) == (0.3, 0.7)
    assert rj.move_probabilities(10, 10, 0.3, False) == (0.3, 1.0)
    assert rj.move_probabilities(9, 10, 0.3, True) == (0.3, 1.0)
    assert rj.move_probabilities(1, 10, 0.3, True) == (1.0, 0.7)
    assert rj.move_probabilities(2, 10, 0.3, False) == (1.0, 0.7)


def test_reference_known_answers_overlap_and_no_overlap():
    """tests/test_reversible_jump.py:347-434 of the reference, restated on the oracle: a knot born on top of an existing
    one splits its coefficient 50/50 and log|F| = log 0.5; a well separated knot leaves the others alone."""
    from oracle import rj

    X = np.linspace(-10, 10, 50)
    m = dict(X=X, y=None, tau_y=1.0, tau_beta=0.25, mu_beta=0.0, rho=8.0, a_omega=3.0, b_omega=2.0, theta_lo=-10.0,
             theta_hi=10.0, n_max=20, birth_probability=0.5, match_scale=1.0, match_limits=(-10.0, 10.0))
    theta, omega, beta = np.array([-10.0, -5.0, 5.0, 10.0]), np.ones(4), np.ones(4)
    st = dict(n=4, theta=theta, omega=omega, beta=beta, B=rj.make_basis(X, theta, omega))
    _, info = rj.rj_step(m, st, dict(u_move=0.1, theta_new=10.0, omega_new=1.0, beta_new=None, u_accept=0.5))
    assert info["birth"]
    np.testing.assert_allclose(info["prop"]["beta"][-2:], [0.5, 0.5], atol=1e-5)
    np.testing.assert_allclose(info["prop"]["beta"].sum(), 4.0, atol=1e-5)
    _, info = rj.rj_step(m, st, dict(u_move=0.1, theta_new=0.0, omega_new=1.0, beta_new=None, u_accept=0.5))
    np.testing.assert_allclose(info["prop"]["beta"], [1, 1, 1, 1, 0], atol=1e-6)
    # death of the last knot when it overlaps its neighbour: the survivor takes both coefficients, log|F| = log 0.5
    theta2 = np.array([-10.0, -5.0, 10.0, 10.0])
    st2 = dict(n=4, theta=theta2, omega=omega, beta=beta, B=rj.make_basis(X, theta2, omega))
    _, info = rj.rj_step(m, st2, dict(u_move=0.9, del_index=3, u_accept=0.5))
    assert not info["birth"]
    np.testing.assert_allclose(info["prop"]["beta"][-1], 2.0, atol=1e-5)
    p_birth, p_death = rj.move_probabilities(4, 20, 0.5, False)
    np.testing.assert_allclose(info["lq_fwd"] - np.log(p_death), np.log(0.5), atol=1e-5)


# ------------------------------------------------------------------------------------------------ companion samplers
MOVE_NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "rjmoves_*.npz")))


def moves_model_of(g):
    return dict(X=g["X"], y=g["y"] if str(g["response"]) == "normal" else None, tau_y=float(g["tau_y"]),
                tau_beta=float(g["tau_beta"]), mu_beta=float(g["mu_beta"]), rho=float(g["rho"]), a_omega=float(g["a_omega"]),
                b_omega=float(g["b_omega"]), theta_lo=float(g["theta_lo"]), theta_hi=float(g["theta_hi"]),
                n_max=int(g["n_max"]))


@pytest.mark.parametrize("name", MOVE_NAMES)
def test_rj_companion_oracle_replays_reference_calls(name):
    """ManifoldMALA on the coefficients and RandomWalkLoop on knots / widths (basis rebuilt per proposal) of the RJ
    model, call by call against the live reference's recorded states and variates."""
    from oracle import rj

    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    m = moves_model_of(g)
    seen = set()
    for it in range(g["kind"].size):
        n, kind = int(g["n"][it]), int(g["kind"][it])
        th, om, be = (g[k + "_before"][it][:n] for k in ("theta", "omega", "beta"))
        st = dict(n=n, theta=th, omega=om, beta=be, B=rj.make_basis(g["X"], th, om))
        if kind == 0:
            new, info = rj.coef_mmala_step(m, st, float(g["step_beta"]), g["z"][it][:n], g["u"][it][0])
            n_acc = int(info["accepted"])
        else:
            which = "theta" if kind == 1 else "omega"
            lim = (m["theta_lo"], m["theta_hi"]) if kind == 1 else (float(g["omega_lo"]), float(g["omega_hi"]))
            new, n_acc = rj.knot_walk_sweep(m, st, which, float(g["step_" + which]), lim, g["tn_u"][it][:n], g["u"][it][:n])
        assert n_acc == int(g["accepted"][it]), (it, kind)
        for k in ("theta", "omega", "beta"):
            np.testing.assert_allclose(new[k], g[k + "_after"][it][:n], rtol=1e-9, atol=1e-10, err_msg=f"{it} {kind} {k}")
        seen.add((kind, n))
    assert len({k for k, _ in seen}) == 3 and len({n for _, n in seen}) >= 2
