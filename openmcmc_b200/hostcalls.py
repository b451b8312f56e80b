"""Host-array entry points: run ONE device operation on a reference-style dict state (numpy in, numpy out).

This is what `sampler.sample(state)`, `dist.log_p(state)`, `model.grad_log_p(...)` and `param.predictor(state)` call
when used exactly like the reference objects (n_chains = 1, host arrays).  Each call uploads what it needs, launches
the same CUDA kernels the resident sweep plan uses, and downloads the result — host<->device copies included, which is
what bench.py's `e2e` leg measures.  There is no CPU arithmetic here.
"""

import itertools

import numpy as np
import torch

from openmcmc_b200 import engine
from openmcmc_b200 import kernels as K

_call_counter = itertools.count(1)
_default_seed = 0


def set_seed(seed: int):
    """Seed for the single-call (`.sample(state)`) RNG stream; the reference uses numpy's global RandomState."""
    global _default_seed, _call_counter
    _default_seed = int(seed)
    _call_counter = itertools.count(1)


def _run_ops(ops):
    for _, fn in ops:
        fn()


def sample_once(sampler, state: dict, debug_draws: dict = None) -> dict:
    """ref: MCMCSampler.sample contract (sampler.py:57-67): returns the state with state[param] replaced."""
    dev = K.init_device()
    from openmcmc_b200 import gmrf_plan
    from openmcmc_b200.model import Model

    user_state, user_model = state, sampler.model
    dists, state, changed = engine.unreplicate(list(sampler.model.values()), state, frozenset({sampler.param}))
    st = engine.DeviceState(1, dev, state, per_chain_names={sampler.param})
    plan = engine.Plan(st, seed=_default_seed, recenter=False)
    plan.sweep_counter.fill_(next(_call_counter))
    try:
        if changed:   # replicated data response: compiled in its single-column form, the caller's objects stay as they are
            sampler.model = Model(dists, response=getattr(user_model, "response", None))
        gmrf_plan.discover(plan, state, list(sampler.model.values()))
        sampler.compile(plan, state, debug_draws)
    finally:
        sampler.model = user_model
    state = user_state
    _run_ops(plan.ops)
    torch.cuda.synchronize()
    status = int(plan.status.max().item())
    if status & 1:
        raise np.linalg.LinAlgError("Matrix is not positive definite")  # reference: np.linalg.cholesky raises
    new = st.get_host(sampler.param)
    old = state[sampler.param]
    state[sampler.param] = new.reshape(np.shape(old)) if np.ndim(old) == 2 else new
    if hasattr(sampler, "trim_host_state") and list(getattr(sampler, "extra_state_names", lambda: [])()):
        # ReversibleJump rewrites knots, widths, coefficients and the basis matrix together with the count: the returned
        # state is consistent, in the reference's shapes (sampler.py:57-67 contract)
        sampler.trim_host_state(state, st.get_host)
    if hasattr(sampler, "_after_sample"):
        sampler._after_sample(plan)
    return state


def linear_predictor(state: dict, terms):
    """sum_t state[X_t] @ state[theta_t]  (ref: parameter.py:174-197) on the device."""
    if not terms:
        return 0
    dev = K.init_device()
    st = engine.DeviceState(1, dev, state)
    n = None
    vt = []
    n_rep = 1
    for xname, tname, *flag in terms:
        X, th = st[xname], st[tname]
        if X.kind != "dense":
            raise engine.PlanError("LinearCombination.predictor with structured prefactors is not supported yet")
        if th.cols != 1:
            raise engine.PlanError("LinearCombination.predictor with replicated parameters is not supported yet")
        n = X.rows
        vt.append((X.vec(), th.vec(), X.cols, bool(flag[0]) if flag else False))
    out = torch.empty(1, n, dtype=torch.float64, device=st.device)
    for i in range(0, len(vt), 4):
        part = out if i == 0 else torch.empty_like(out)
        K.linear_predictor(1, n, vt[i:i + 4], part)
        if i:
            raise engine.PlanError("LinearCombination with more than 4 terms is not supported yet")
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(n, n_rep)


def scaled_matrix(state: dict, matrix: str, scalar: str):
    """scalar * matrix in the format of the matrix (ref: parameter.py:319-329; the scalar must hold one value, as the
    reference's `.item()` demands).  The values are scaled on the device (omc_combine)."""
    from scipy import sparse

    sc = np.asarray(state[scalar], dtype=np.float64)
    if sc.size != 1:
        raise ValueError("can only convert an array of size 1 to a Python scalar")
    m = state[matrix]
    dev = torch.device("cuda", K.init_device())
    vals = m.data if sparse.issparse(m) else np.asarray(m, dtype=np.float64)
    if vals.size == 0:
        return m.copy()
    x = K.upload(np.ascontiguousarray(vals, dtype=np.float64).reshape(-1), dev)
    s_dev = K.upload(sc.reshape(1), dev)
    out = torch.empty_like(x)
    K.combine(1, x.numel(), [K.vec(x)], [K.vec(s_dev)], out)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    if sparse.issparse(m):
        r = m.astype(np.float64).copy()
        r.data = res
        return r
    return res.reshape(np.shape(m))


def scale_columns(X, theta):
    """X * exp(theta)' (column j scaled by exp(theta_j)), the Jacobian factor of LinearCombinationWithTransform.grad
    (ref: parameter.py:282-297): exp and the products run in omc_linear_predictor, one column at a time."""
    from scipy import sparse

    Xd = X.toarray() if sparse.issparse(X) else np.asarray(X, dtype=np.float64)
    n, p = Xd.shape
    dev = torch.device("cuda", K.init_device())
    th = K.upload(np.asarray(theta, dtype=np.float64).reshape(-1), dev)
    out = torch.empty(p, n, dtype=torch.float64, device=dev)
    cols = K.upload(np.ascontiguousarray(Xd.T), dev)              # [p, n]: "chain" j = column j, a 1-term predictor each
    K.linear_predictor(p, n, [(K.vec(cols, n), K.vec(th, 1), 1, True)], out)
    torch.cuda.synchronize()
    return out.cpu().numpy().T


def log_p(dist, state: dict, by_observation: bool = False):
    from openmcmc_b200 import devdist

    return devdist.log_p_host(dist, state, by_observation)


def grad_log_p(dist, state: dict, param: str, hessian_required: bool, method: str):
    from openmcmc_b200 import devdist

    return devdist.grad_log_p_host(dist, state, param, hessian_required, method)


def rvs(dist, state: dict, n: int = 1):
    from openmcmc_b200 import devdist

    return devdist.rvs_host(dist, state, n)
