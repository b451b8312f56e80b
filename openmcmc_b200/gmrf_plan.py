"""Plan fragments for the temporal-GMRF path (BASELINE configs[2]): a field b with a tridiagonal prior precision
lambda*P and a Normal likelihood y ~ N(b, (tau*W)^-1), W diagonal.  ref: sampler.py:154-207 with gmrf.py:167-198 on the
sparse branch (gmrf.py:489-520); quadratic forms for the NormalGamma updates sampler.py:275-284.

The reference picks dense or sparse linear algebra from the *Python types* in the state (example 4 as written
runs dense); here the structure of the model decides: tridiagonal prior + diagonal likelihood => omc_tridiag_nn_draw.
Host logic only: which state entries feed which kernel argument, and which derived quantities a launch refreshes.
"""

import numpy as np
import torch
from scipy import sparse

from openmcmc_b200 import engine
from openmcmc_b200 import kernels as K
from openmcmc_b200.parameter import Identity, LinearCombination


class GMRFField:
    """One latent field and the two Normal distributions that meet in its conditional: owner of the tridiagonal
    workspace, of the quadratic forms (x-mu0)'P(x-mu0) and (y-x)'W(y-x), and of log|P|."""

    def __init__(self, plan: "engine.Plan", name: str):
        st = plan.state
        self.plan = plan
        self.name = name
        self.x = st[name]
        if not self.x.per_chain:
            self.x = st.put(name, self.x.data, per_chain=True)
        if self.x.cols != 1:
            raise engine.PlanError("replicated GMRF fields (n_rep > 1) are not supported by the device path")
        self.n = self.x.rows
        C = st.n_chains
        self.prior = None   # dict(nrm, P, pd, pe, lam, mu0, h, cnt)
        self.lik = None     # dict(nrm, y, W, tau, cnt)
        self.ss_prior = plan.new(C, fill=0.0)
        self.ss_lik = plan.new(C, fill=0.0)
        ws_bytes = K.tridiag_workspace(C, self.n)
        self.ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=st.device)
        plan.keep.append(self.ws)
        self._logdet_P = None
        self.q_prior = self.q_lik = None

    # ---------------------------------------------------------------- slots
    def set_prior(self, host_state, nrm):
        if self.prior is not None:
            return
        plan, st = self.plan, self.plan.state
        if not isinstance(nrm.mean, Identity):
            raise engine.PlanError("GMRF prior mean must be an Identity parameter")
        mname, sname = engine._scalar_and_matrix(nrm.precision)
        P = engine.ensure_matrix(st, host_state, mname)
        if P.kind == "dense" or P.per_chain:
            raise engine.PlanError("GMRF prior precision must be tridiagonal / diagonal and shared by all chains")
        dev = st.device
        if P.kind == "eye":
            pd, pe = torch.ones(self.n, dtype=torch.float64, device=dev), None
        elif P.kind == "diag":
            pd, pe = P.data.reshape(-1), None
        else:
            pd, pe = P.data.reshape(-1), P.off.reshape(-1)
        plan.keep.extend([t for t in (pd, pe) if t is not None])
        mu0 = st[nrm.mean.form]
        mu_host = host_state.get(nrm.mean.form)
        mu_zero = mu_host is not None and not isinstance(mu_host, torch.Tensor) and not np.any(np.asarray(mu_host))
        h = None
        if mu_zero:
            mu0 = None
        else:
            Cm = st.n_chains if mu0.per_chain else 1
            h = plan.new(Cm, self.n)
            K.tridiag_matvec(pd, pe, mu0.vec(), Cm, self.n, h)   # constant data: once at plan time
        Ph = host_state.get(mname)
        logdet_key = None
        if sparse.issparse(Ph):     # identity of the constant host matrix (and a checksum of its values) for the log|P| cache
            logdet_key = (id(Ph), Ph.shape, Ph.nnz, float(Ph.data.sum()) if Ph.nnz else 0.0)
        self.prior = dict(nrm=nrm, P=P, pd=pd, pe=pe, lam=st[sname] if sname else None, mu0=mu0, h=h, logdet_key=logdet_key,
                          h_per_chain=bool(mu0 is not None and mu0.per_chain), cnt=float(P.npos),
                          deps=frozenset({self.name, nrm.mean.form, mname}))
        self.q_prior = f"quad[{nrm.response}]"
        self._register()

    def set_lik(self, host_state, nrm):
        if self.lik is not None:
            return
        st = self.plan.state
        mname, sname = engine._scalar_and_matrix(nrm.precision)
        W = engine.ensure_matrix(st, host_state, mname)
        if W.kind not in ("eye", "diag"):
            raise engine.PlanError("the likelihood precision of a GMRF field must be (a scalar times) a diagonal matrix")
        y = st[nrm.response]
        if y.size != self.n or y.cols != 1:
            raise engine.PlanError(f"response '{nrm.response}' does not match the GMRF field '{self.name}'")
        self.lik = dict(nrm=nrm, y=y, W=W, tau=st[sname] if sname else None, cnt=float(W.npos),
                        deps=frozenset({self.name, nrm.response, mname}))
        self.q_lik = f"quad[{nrm.response}]"
        self._register()

    def _register(self):
        """(Re-)declare the quadratic-form quantities; one omc_tridiag_quadforms launch refreshes both."""
        plan = self.plan
        if self.prior is not None:
            sib = (self.q_lik,) if self.lik is not None else ()
            plan.add_quantity(engine.Quantity(self.q_prior, self.prior["deps"], self._emit_quadforms, sib))
        if self.lik is not None:
            sib = (self.q_prior,) if self.prior is not None else ()
            plan.add_quantity(engine.Quantity(self.q_lik, self.lik["deps"], self._emit_quadforms, sib))

    # ---------------------------------------------------------------- kernel arguments
    def args(self, **extra):
        C, n = self.plan.state.n_chains, self.n
        pr, lk = self.prior, self.lik
        dev = self.plan.state.device
        if pr is None:   # likelihood-only use (long diagonal quadratic form): P = 0
            if not hasattr(self, "_zero_pd"):
                self._zero_pd = torch.zeros(n, dtype=torch.float64, device=dev)
            pd, pe = self._zero_pd, None
        else:
            pd, pe = pr["pd"], pr["pe"]
        kw = dict(lam=pr["lam"].vec() if pr and pr["lam"] is not None else None,
                  mu0=pr["mu0"].vec() if pr and pr["mu0"] is not None else None,
                  h=(K.vec(pr["h"], n if pr["h_per_chain"] else None) if pr and pr["h"] is not None else None))
        if lk is not None:
            kw.update(tau=lk["tau"].vec() if lk["tau"] is not None else None, y=lk["y"].vec(),
                      w=lk["W"].vec() if lk["W"].kind == "diag" else None)
        kw.update(extra)
        return K.tridiag_args(C, n, pd, pe, self.ws, **kw)

    def _emit_quadforms(self):
        def launch():
            K.tridiag_quadforms(self.args(x=self.x.data, ss_prior=self.ss_prior, ss_lik=self.ss_lik))

        self.plan.emit(launch, f"tridiag_quadforms[{self.name}]")

    def logdet_P(self):
        """log|P| of the constant prior precision: one factorisation-only launch at plan time (lambda = 1, tau = 0)."""
        if self._logdet_P is None:
            plan, pr = self.plan, self.prior
            dev = plan.state.device
            out = plan.new(1)
            key = pr.get("logdet_key")
            if key is not None and key in _LOGDET_CACHE:      # the same constant matrix in an earlier plan of this process
                out.fill_(_LOGDET_CACHE[key])
                self._logdet_P = out
                return out
            zero = torch.zeros(1, dtype=torch.float64, device=dev)
            ws = torch.zeros(K.tridiag_workspace(1, self.n), dtype=torch.uint8, device=dev)
            K.tridiag_nn_draw(K.tridiag_args(1, self.n, pr["pd"], pr["pe"], ws, tau=K.vec(zero), y=K.vec(pr["pd"]),
                                             logdet=out))
            torch.cuda.current_stream().synchronize()
            if key is not None:
                _LOGDET_CACHE[key] = float(out.item())
            self._logdet_P = out
        return self._logdet_P


_LOGDET_CACHE = {}     # logdet_key of a constant prior precision -> log|P| (one factorisation per matrix and process)


def _fields(plan):
    return plan.__dict__.setdefault("_gmrf_fields", {})


def get_field(plan, name) -> GMRFField:
    f = _fields(plan)
    if name not in f:
        f[name] = GMRFField(plan, name)
    return f[name]


def _identity_mean_of(plan, host_state, nrm):
    """Name of the field a Normal's mean equals: Identity(name), or LinearCombination({name: I}) with I an identity
    matrix (the sparse-enabled way of writing example 4).  None otherwise."""
    if isinstance(nrm.mean, Identity):
        return nrm.mean.form
    if isinstance(nrm.mean, LinearCombination) and len(nrm.mean.form) == 1:
        (prm, pref), = nrm.mean.form.items()
        arr = plan.state.arrays.get(pref)
        if arr is not None:
            return prm if arr.kind == "eye" else None
        shape = getattr(host_state.get(pref), "shape", None)
        if shape is None or len(shape) != 2 or shape[0] != shape[1]:
            return None     # a rectangular design matrix: the regression path
        if not sparse.issparse(host_state[pref]) and shape[0] > 64:
            return None     # a large dense square prefactor is data, not an identity
        if engine.classify_matrix(host_state[pref])[0] == "eye":
            engine.ensure_matrix(plan.state, host_state, pref)
            return prm
    return None


def discover(plan, host_state, dists):
    """Register every GMRF field in `dists`: a Normal whose un-scaled precision is tridiagonal defines a field (prior
    slot); a Normal whose mean is that field and whose precision is diagonal fills the likelihood slot."""
    from openmcmc_b200.distribution.location_scale import Normal

    normals = [d for d in dists if type(d) is Normal]
    for d in normals:
        try:
            mname, _ = engine._scalar_and_matrix(d.precision)
        except engine.PlanError:
            continue
        if mname not in host_state:
            continue
        P = engine.ensure_matrix(plan.state, host_state, mname)
        if P.kind == "tridiag" and not P.per_chain:
            get_field(plan, d.response).set_prior(host_state, d)
    for d in normals:
        if d.response in _fields(plan):
            continue
        name = _identity_mean_of(plan, host_state, d) if _quick_mean_name(d) in _fields(plan) else None
        if name in _fields(plan):
            get_field(plan, name).set_lik(host_state, d)


def _quick_mean_name(nrm):
    if isinstance(nrm.mean, Identity):
        return nrm.mean.form
    if isinstance(nrm.mean, LinearCombination) and len(nrm.mean.form) == 1:
        return next(iter(nrm.mean.form))
    return None


def is_gmrf_update(plan, host_state, param, prior, lik) -> bool:
    """NormalNormal(param) goes down the tridiagonal path when its prior precision is tridiagonal (or the field is too
    long for the dense kernel) and the likelihood mean is the field itself."""
    mname, _ = engine._scalar_and_matrix(prior.precision)
    P = engine.ensure_matrix(plan.state, host_state, mname)
    if P.kind == "dense":
        return False
    if _quick_mean_name(lik) != param or _identity_mean_of(plan, host_state, lik) != param:
        return False
    return P.kind == "tridiag" or plan.state[param].rows > 64


def compile_normal_normal_identity(sampler, plan, host_state, prior, lik, debug_draws):
    """NormalNormal.sample for a GMRF field.  ref: sampler.py:154-207 -> gmrf.py:167-198 (sparse branch)."""
    st = plan.state
    C = st.n_chains
    fld = get_field(plan, sampler.param)
    fld.set_prior(host_state, prior)
    fld.set_lik(host_state, lik)
    n = fld.n
    ctx = plan.ctx(sampler)
    if "rng" not in ctx:
        ctx["rng"] = plan.rng_site()
        ctx["dz"], ctx["dz_stride"] = (None, 0)
        if debug_draws and "z" in debug_draws:
            ctx["dz"], ctx["dz_stride"] = plan.debug_tensor(debug_draws["z"], n)
        ctx["probes"] = None
        if plan.probes is not None and plan.probes.get("enable"):
            ctx["probes"] = {"l": plan.new(C, n), "c": plan.new(C, max(n - 1, 1)), "logdet": plan.new(C)}
            plan.probes[sampler.param] = ctx["probes"]
    pr = ctx["probes"] or {}

    def launch():
        K.tridiag_nn_draw(fld.args(x=fld.x.data, rng_=ctx["rng"], debug_z=ctx["dz"], debug_sweep_stride=ctx["dz_stride"],
                                   ss_prior=fld.ss_prior, ss_lik=fld.ss_lik, status=plan.status, probe_l=pr.get("l"),
                                   probe_c=pr.get("c"), logdet=pr.get("logdet")))

    plan.emit(launch, f"tridiag_nn_draw[{sampler.param}]")
    plan.wrote(sampler.param)
    # the draw kernel's epilogue already holds both quadratic forms of the NEW field
    plan.valid[fld.q_prior] = True
    plan.valid[fld.q_lik] = True


def long_quadratic_form(plan, host_state, nrm, P, x, mu):
    """Quadratic form of a long Normal (tridiagonal precision, or a diagonal one too long for omc_quadform).
    Returns (ss_vec_fn, cnt_vec_fn, quantity_name) like engine.get_quadratic_form."""
    fields = _fields(plan)
    mean_name = _quick_mean_name(nrm)
    if nrm.response in fields or P.kind == "tridiag":
        fld = get_field(plan, nrm.response)
        fld.set_prior(host_state, nrm)
        cnt = plan.new(plan.state.n_chains, fill=fld.prior["cnt"])
        return (lambda: K.vec(fld.ss_prior, 1)), (lambda: K.vec(cnt, 1)), fld.q_prior
    if mean_name in fields and _identity_mean_of(plan, host_state, nrm) == mean_name:
        fld = fields[mean_name]
        fld.set_lik(host_state, nrm)
        cnt = plan.new(plan.state.n_chains, fill=fld.lik["cnt"])
        return (lambda: K.vec(fld.ss_lik, 1)), (lambda: K.vec(cnt, 1)), fld.q_lik
    # a long diagonal Normal on its own: treat the response as a field with a diagonal "prior"
    fld = get_field(plan, nrm.response)
    fld.set_prior(host_state, nrm)
    cnt = plan.new(plan.state.n_chains, fill=fld.prior["cnt"])
    return (lambda: K.vec(fld.ss_prior, 1)), (lambda: K.vec(cnt, 1)), fld.q_prior


def tridiag_logdet(plan, P, out):
    """log|P| for a tridiagonal constant matrix (engine.logdet_of)."""
    for fld in _fields(plan).values():
        if fld.prior is not None and fld.prior["P"] is P:
            out.copy_(fld.logdet_P())
            return
    raise engine.PlanError("log-determinant of a tridiagonal matrix that is not a registered GMRF prior precision")
