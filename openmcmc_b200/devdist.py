"""Distributions on the device: pattern-matching of a conditional Model onto the term list of the MH kernels
(include/omc.h: omc_mh_model_t) and the host-array entry points behind `dist.log_p`, `dist.grad_log_p`, `dist.rvs`.

ref: model.py:57-112 (Model.log_p / grad_log_p fan-out), distribution.py:241-261, 490-508, 422-442,
location_scale.py:145-250.  Host logic only (shapes and matching); every value is computed by libomc kernels.
"""

import numpy as np
import torch
from scipy import sparse

from openmcmc_b200 import engine
from openmcmc_b200 import kernels as K
from openmcmc_b200.parameter import Identity

F64 = torch.float64


def _expand(plan, arr: "engine.DevArray", p_dim: int, n_rep: int):
    """Operand of a term: returns (vec, len) with len in {1, n_elem}; (p_dim,1) operands of a replicated parameter are
    broadcast along the replicate axis once at plan time (constant data)."""
    n_elem = p_dim * n_rep
    if arr.size in (1, n_elem):
        return arr.vec(), arr.size
    if arr.size == p_dim and arr.cols == 1:
        t = arr.data.reshape(-1, p_dim, 1).expand(-1, p_dim, n_rep).contiguous()
        plan.keep.append(t)
        return (K.vec(t, n_elem) if arr.per_chain else K.vec(t)), n_elem
    raise engine.PlanError(f"operand of shape ({arr.rows},{arr.cols}) cannot be broadcast to ({p_dim},{n_rep})")


def build_terms(plan: "engine.Plan", host_state: dict, model, param: str):
    """Term list of the conditional model of `param` (every member distribution of `model` that depends on it).

    Returns (omc_mh_model_t, labels).  Raises PlanError for combinations the device path does not implement.
    """
    from openmcmc_b200.distribution.distribution import Gamma, Poisson, Uniform
    from openmcmc_b200.distribution.location_scale import LogNormal, Normal, NullDistribution
    from openmcmc_b200.parameter import LinearCombination

    st = plan.state
    theta = st[param]
    p_dim, n_rep = theta.rows, theta.cols
    n_elem = p_dim * n_rep
    terms, labels = [], []
    for dist in model.values():
        if param not in dist.param_list or isinstance(dist, NullDistribution):
            continue
        if isinstance(dist, LogNormal) and dist.response != param:
            # mean-parameter branch (location_scale.py:344-347, 401-404): the Normal of log(y); the Jacobian term
            # -sum(log y) does not depend on theta
            dist, _ = engine.lognormal_shim(plan, dist)
        if isinstance(dist, Poisson) and isinstance(dist.rate, Identity) and dist.rate.form == param:
            k = st[dist.response]
            if k.size != n_elem:
                raise engine.PlanError(f"Poisson response '{dist.response}' does not match the shape of '{param}'")
            terms.append(K.term(K.TERM_POISSON_RATE, data=k.vec()))
        elif isinstance(dist, Gamma) and dist.response == param:
            if not isinstance(dist.shape, Identity) or not isinstance(dist.rate, Identity):
                raise engine.PlanError("MH on a Gamma response needs Identity shape and rate parameters")
            (v1, l1), (v2, l2) = _expand(plan, st[dist.shape.form], p_dim, n_rep), _expand(plan, st[dist.rate.form],
                                                                                         p_dim, n_rep)
            terms.append(K.term(K.TERM_GAMMA_RESPONSE, p1=v1, p1_len=l1, p2=v2, p2_len=l2))
        elif isinstance(dist, LogNormal) and dist.response == param:
            # location_scale.py:296-303 (log_p), :340-343 / :383-399 (response-branch gradient and Hessian)
            if not isinstance(dist.mean, Identity):
                raise engine.PlanError("MH on a LogNormal response needs an Identity mean parameter")
            mname, sname = engine._scalar_and_matrix(dist.precision)
            P = engine.ensure_matrix(st, host_state, mname)
            if P.kind == "tridiag" or (n_rep != 1 and P.kind != "eye"):
                raise engine.PlanError("MH on a LogNormal response supports eye / diagonal / dense precisions, n_rep = 1")
            v1, l1 = _expand(plan, st[dist.mean.form], p_dim, n_rep)
            logdet = engine.logdet_of(plan, P)
            terms.append(K.term(K.TERM_LOGNORMAL_RESPONSE, mat_kind=engine._mat_kind(P), p1=v1, p1_len=l1, P=P.vec(),
                                scalar=st[sname].vec() if sname else None,
                                logdet=K.vec(logdet, 1 if P.per_chain else None) if logdet is not None else None))
        elif (isinstance(dist, Normal) and isinstance(dist.mean, LinearCombination) and param in dist.mean.form
              and dist.response != param):
            # theta enters the mean of a Normal response linearly (optionally through exp): location_scale.py:234-250 with
            # parameter.py:199-228 / 283-297.  Evaluated from the data-only regression record (omc.h: NORMAL_LINEAR).
            if n_rep != 1:
                raise engine.PlanError("MH through a linear Normal mean needs a column-vector parameter")
            rl = engine.get_regression(plan, host_state, dist, param, data_only=True)
            plan.require(rl.q_gg)
            logdet = engine.logdet_of(plan, rl.W)
            transform = bool((getattr(dist.mean, "transform", None) or {}).get(param, False))
            terms.append(K.term(K.TERM_NORMAL_LINEAR, stats=K.vec(rl.stats, rl.rec), n_data=rl.n,
                                scalar=st[rl.scalar].vec() if rl.scalar else None,
                                logdet=K.vec(logdet, 1 if rl.W.per_chain else None) if logdet is not None else None,
                                transform_exp=transform))
        elif isinstance(dist, Normal) and dist.response == param:
            if not isinstance(dist.mean, Identity):
                raise engine.PlanError("MH on a Normal response needs an Identity mean parameter")
            mname, sname = engine._scalar_and_matrix(dist.precision)
            P = engine.ensure_matrix(st, host_state, mname)
            if P.kind == "tridiag":
                raise engine.PlanError("MH with a tridiagonal Normal prior is not supported by the device path")
            if n_rep != 1 and P.kind != "eye":
                raise engine.PlanError("MH on a replicated Normal response supports identity precision matrices only")
            v1, l1 = _expand(plan, st[dist.mean.form], p_dim, n_rep)
            logdet = engine.logdet_of(plan, P)
            lo = -np.inf if dist.domain_response_lower is None else float(np.max(dist.domain_response_lower))
            hi = np.inf if dist.domain_response_upper is None else float(np.min(dist.domain_response_upper))
            for lim in (dist.domain_response_lower, dist.domain_response_upper):
                if lim is not None and np.ptp(np.asarray(lim, dtype=float)) != 0:
                    raise engine.PlanError("element-wise different Normal domain limits are not supported yet")
            terms.append(K.term(K.TERM_NORMAL_RESPONSE, mat_kind=engine._mat_kind(P), p1=v1, p1_len=l1, P=P.vec(),
                                scalar=st[sname].vec() if sname else None,
                                logdet=K.vec(logdet, 1 if P.per_chain else None) if logdet is not None else None,
                                dom_lo=lo, dom_hi=hi))
        elif (isinstance(dist, Normal) and isinstance(dist.mean, Identity) and dist.mean.form == param
              and dist.response != param):
            # theta is the (Identity) mean of another Normal response y: (y-theta)'Q(y-theta) is the response form with
            # the roles of y and theta swapped; grad = Q(y-theta), H = Q  (location_scale.py:234-242 with grad = eye)
            mname, sname = engine._scalar_and_matrix(dist.precision)
            P = engine.ensure_matrix(st, host_state, mname)
            if P.kind == "tridiag" or (n_rep != 1 and P.kind != "eye"):
                raise engine.PlanError("MH through a Normal mean supports eye / diagonal / dense precisions, n_rep = 1")
            yv, yl = _expand(plan, st[dist.response], p_dim, n_rep)
            logdet = engine.logdet_of(plan, P)
            terms.append(K.term(K.TERM_NORMAL_RESPONSE, mat_kind=engine._mat_kind(P), p1=yv, p1_len=yl, P=P.vec(),
                                scalar=st[sname].vec() if sname else None,
                                logdet=K.vec(logdet, 1 if P.per_chain else None) if logdet is not None else None))
        elif isinstance(dist, Uniform) and dist.response == param:
            lo = st.put(f"__uniform_lo[{param}]", np.broadcast_to(dist.domain_response_lower, (p_dim, 1)).copy())
            hi = st.put(f"__uniform_hi[{param}]", np.broadcast_to(dist.domain_response_upper, (p_dim, 1)).copy())
            (v1, l1), (v2, l2) = _expand(plan, lo, p_dim, n_rep), _expand(plan, hi, p_dim, n_rep)
            terms.append(K.term(K.TERM_UNIFORM_RESPONSE, p1=v1, p1_len=l1, p2=v2, p2_len=l2))
        else:
            raise engine.PlanError(
                f"the device MH path has no term for {type(dist).__name__}('{dist.response}') as a function of '{param}'")
        labels.append(f"{type(dist).__name__}[{dist.response}]")
    return K.mh_model(st.n_chains, n_elem, terms), labels


# ============================================================================================== host-array calls
def _one_chain_plan(state: dict, per_chain=()):
    dev = K.init_device()
    st = engine.DeviceState(1, dev, state, per_chain_names=set(per_chain))
    return engine.Plan(st), st


def log_p_host(dist, state: dict, by_observation: bool = False):
    """dist.log_p(state) for a host dict state (n_chains = 1).  ref: distribution.py:40-54"""
    from openmcmc_b200.model import Model

    if by_observation:
        return _log_p_by_observation(dist, state)
    (dist,), state, _ = engine.unreplicate([dist], state)
    plan, st = _one_chain_plan(state)
    out = plan.new(1)
    plan.ops = []
    engine.compile_log_post(plan, state, Model([dist]), out)
    for _, fn in plan.ops:
        fn()
    torch.cuda.synchronize()
    return float(out.item())


def _log_p_by_observation(dist, state: dict):
    """log_p(state, by_observation=True): one value per COLUMN of the response (its replicates), rows summed.
    ref: distribution.py:241-261 (Gamma), 422-442 (Uniform), 490-508 (Poisson); location_scale.py:145-167 ->
    gmrf.py:321-348 (Normal).  Every column is handed to the log-density kernels as a "chain" of its own."""
    from openmcmc_b200 import gmrf
    from openmcmc_b200.distribution.distribution import Gamma, Poisson, Uniform
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.parameter import Identity

    x = np.asarray(state[dist.response], dtype=np.float64)
    x = x.reshape(-1, 1) if x.ndim < 2 else x
    p, n = x.shape
    if isinstance(dist, Uniform):
        rng_ = np.broadcast_to(dist.domain_response_upper - dist.domain_response_lower, (p, 1))
        return np.ones(n) * -float(np.sum(np.log(rng_)))
    if type(dist) is Normal:
        lo, hi = dist.domain_response_lower, dist.domain_response_upper      # location_scale.py:164-165, 169-188
        if (lo is not None and np.any(x < lo)) or (hi is not None and np.any(x > hi)):
            return -np.inf
        return gmrf.multivariate_normal_pdf(x, dist.mean.predictor(state), dist.precision.predictor(state), by_observation=True)
    if isinstance(dist, (Gamma, Poisson)):
        dev = torch.device("cuda", K.init_device())

        def cols(par):      # [p or 1, n or 1] parameter -> per-column device rows [n, p'] (layout only)
            if not isinstance(par, Identity):
                raise engine.PlanError("log_p(by_observation=True): Identity shape / rate parameters only")
            a = np.asarray(state[par.form], dtype=np.float64)
            a = a.reshape(-1, 1) if a.ndim < 2 else a
            if a.shape[0] not in (1, p) or a.shape[1] not in (1, n):
                raise ValueError(f"parameter '{par.form}' of shape {a.shape} for a response of shape {(p, n)}")
            t = K.upload(np.ascontiguousarray(a.T), dev)            # [n or 1, p or 1]
            return t, (a.shape[0] if a.shape[1] == n else 0), a.shape[0]

        xd = K.upload(np.ascontiguousarray(x.T), dev)               # [n, p]
        out = torch.empty(n, dtype=torch.float64, device=dev)
        if isinstance(dist, Gamma):
            (sh, sh_stride, sh_len), (rt, rt_stride, rt_len) = cols(dist.shape), cols(dist.rate)
            K.logp_gamma(n, p, K.vec(xd, p), K.vec((sh, sh_stride)), sh_len, K.vec((rt, rt_stride)), rt_len, out, 0)
        else:
            rt, rt_stride, rt_len = cols(dist.rate)
            K.logp_poisson(n, p, K.vec(xd, p), K.vec((rt, rt_stride)), rt_len, out, 0)
        torch.cuda.synchronize()
        return out.cpu().numpy()
    raise engine.PlanError(f"log_p(by_observation=True) is not provided for {type(dist).__name__} on the device")


def grad_log_p_host(dist, state: dict, param: str, hessian_required: bool, method: str):
    """dist.grad_log_p(state, param): gradient of +log p with the shape of state[param] and Hessian of -log p (d x d).

    ref: distribution.py:90-198 (finite differences), location_scale.py:190-250 (analytic Normal).
    """
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination

    if method not in ("fd", "analytic"):
        raise ValueError("method must be 'fd' (the reference's finite differences) or 'analytic'")
    (dist,), state, _ = engine.unreplicate([dist], state, sampled=frozenset({param}))
    plan, st = _one_chain_plan(state, per_chain=(param,))
    shape = np.shape(state[param])
    if (isinstance(dist, Normal) and isinstance(dist.mean, LinearCombination) and param in dist.mean.form
            and not (getattr(dist.mean, "transform", None) or {}).get(param, False)):
        # linear-mean branch (location_scale.py:234-242): H = scalar * X'WX, g = scalar * X'W(y - X beta) from the fused
        # regression pass
        rl = engine.get_regression(plan, state, dist, param)
        plan.ops = []
        rl._emit_pass()
        for _, fn in plan.ops:
            fn()
        p = rl.p
        tau = float(np.asarray(state[rl.scalar]).item()) if rl.scalar else 1.0
        rec = rl.stats[0]                                    # device record G | g | rss | cnt of the one chain
        Gb = plan.new(1, p)                                  # G beta: the record's G as a p x p design (omc_linear_predictor)
        K.linear_predictor(1, p, [(K.vec(rec[: p * p], 0), K.vec(st[param].data, 0), p)], Gb)
        taus = torch.tensor([tau, -tau], dtype=torch.float64, device=rec.device)
        out = plan.new(1, p + p * p)
        K.combine(1, p, [K.vec(rec[p * p: p * p + p], 0), K.vec(Gb, 0)], [K.vec(taus[0:1], 0), K.vec(taus[1:2], 0)], out[:, :p])
        K.combine(1, p * p, [K.vec(rec[: p * p], 0)], [K.vec(taus[0:1], 0)], out[:, p:])
        torch.cuda.synchronize()
        h = out.cpu().numpy().ravel()
        grad = h[:p].reshape(shape)
        return (grad, h[p:].reshape(p, p)) if hessian_required else grad
    plan.ops = []
    model, _ = build_terms(plan, state, Model([dist]), param)
    for _, fn in plan.ops:      # derived quantities the terms read (e.g. the data-only regression record)
        fn()
    theta = st[param].data
    n = model.n_elem
    grad = plan.new(1, n)
    hess = plan.new(1, n, n, fill=0.0) if hessian_required else None
    K.mh_grad_hess(model, theta, 1 if method == "fd" else 0, grad, hess)
    torch.cuda.synchronize()
    g = grad.cpu().numpy().reshape(shape)
    if hessian_required:
        return g, hess.cpu().numpy().reshape(n, n)
    return g


_rvs_calls = 0


def rvs_host(dist, state: dict, n: int = 1):
    """dist.rvs(state, n): p x n draws on the device -- the prior draws of parameters missing from the initial state
    (mcmc.py:78-80) and the reference's stand-alone use.  Start-up only, never inside a sweep: Normal goes through the
    Alg. 2.5 kernels (`gmrf.sample_normal`), Gamma through `omc_ng_draw` (Philox, Marsaglia-Tsang); Poisson, Uniform
    and Categorical use torch's device generators as glue.  Parity unpinned (free-running variates).
    ref: location_scale.py:252-272, distribution.py:263-278, 354-374, 444-458, 510-523"""
    global _rvs_calls
    from openmcmc_b200 import gmrf as G
    from openmcmc_b200 import hostcalls
    from openmcmc_b200.distribution.distribution import Categorical, Gamma, Poisson, Uniform
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.parameter import Identity, ScaledMatrix

    _rvs_calls += 1
    seed = hostcalls._default_seed * 1_000_003 + _rvs_calls
    dev = torch.device("cuda", K.init_device())
    col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1, 1) if np.ndim(v) < 2 else np.asarray(v, dtype=np.float64)
    if type(dist) is Normal:
        if dist.domain_response_lower is not None or dist.domain_response_upper is not None:
            raise engine.PlanError("rvs of a truncated Normal is not provided: give an initial value in the state")
        mean = col(dist.mean.predictor(state))
        if isinstance(dist.precision, ScaledMatrix):
            Q = state[dist.precision.matrix] * float(np.asarray(state[dist.precision.scalar]).item())
        elif isinstance(dist.precision, Identity):
            Q = state[dist.precision.form]
        else:
            raise engine.PlanError("rvs of a Normal with a mixture precision is not provided")
        Q = Q if sparse.issparse(Q) else np.atleast_2d(np.asarray(Q, dtype=np.float64))
        return G.sample_normal(mean[:, :1], Q=Q, n=n, seed=seed)
    if type(dist) is Gamma:
        shape, rate = col(dist.shape.predictor(state)), col(dist.rate.predictor(state))
        p = max(shape.shape[0], rate.shape[0])
        a0 = torch.as_tensor(shape.reshape(-1)).to(dev)
        b0 = torch.as_tensor(rate.reshape(-1)).to(dev)
        out = torch.empty(n, p, dtype=torch.float64, device=dev)
        sweep = torch.full((1,), _rvs_calls, dtype=torch.int64, device=dev)
        K.ng_draw(n, K.vec(a0, 0), K.vec(b0, 0), K.vec(None), K.vec(None), out, K.rng(seed=seed, sweep=sweep, site=1),
                  n_elem=p, a0_len=a0.numel(), b0_len=b0.numel())
        torch.cuda.synchronize()
        return out.cpu().numpy().T.copy()
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    if type(dist) is Poisson:
        rate = torch.as_tensor(col(dist.rate.predictor(state))).to(dev)
        return torch.poisson(rate.expand(rate.shape[0], n).contiguous(), generator=gen).cpu().numpy()
    if type(dist) is Uniform:
        lo, hi = dist.domain_response_lower, dist.domain_response_upper
        p = np.shape(state[dist.response])[0] if dist.response in state else max(lo.shape[0], hi.shape[0])
        u = torch.rand(p, n, dtype=torch.float64, device=dev, generator=gen).cpu().numpy()
        return lo + (hi - lo) * u
    if type(dist) is Categorical:
        prob = torch.as_tensor(col(dist.prob.predictor(state))).to(dev)
        return torch.multinomial(prob, n, replacement=True, generator=gen).to(torch.float64).cpu().numpy()
    raise engine.PlanError(f"{type(dist).__name__}.rvs on the device is not provided: give an initial value for "
                           f"'{dist.response}' in the state")
