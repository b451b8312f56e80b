"""ReversibleJump sampler (host-side mirror).  ref: sampler/reversible_jump.py:24-373

Same constructor as the reference (`param, model, associated_params, n_max, birth_probability, state_birth_function,
state_death_function, matching_params`) plus one declarative extra, `basis=GaussianKernelBasis(...)`.  The reference's
birth / death callbacks are Python callables over the dict state and cannot run inside a kernel; the
callbacks of the reference's own RJ model (tests/test_reversible_jump.py:62-122: rebuild the Gaussian-kernel basis,
grow / shrink the allocation vector) are what `GaussianKernelBasis` states declaratively.  Giving callables without a
`basis` raises PlanError at compile time.

Device state: fixed capacity `n_max` (padding) — knots, widths, coefficients `[C, n_max]`, basis `[C, n_data, n_max]`,
count `[C]`; one `omc_reversible_jump` launch per sweep does the move choice, proposal, matched coefficient transition,
accept / reject and the in-place state update for every chain (csrc/rj.cu).
"""

from dataclasses import dataclass
from typing import Callable, Union

import numpy as np

from openmcmc_b200 import engine
from openmcmc_b200 import kernels as K
from openmcmc_b200.sampler.metropolis_hastings import MetropolisHastings


# Keep S = B'B of every chain's live basis in the chain state and update it per birth / death (omc.h: omc_rj_t.gram).
# False recomputes it in every step, the first round's form; tests compare the two.
LIVE_GRAM = True


@dataclass
class GaussianKernelBasis:
    """matrix[:, j] = Normal pdf(locations; knots[j], widths[j])  (make_basis of the reference's RJ tests)."""

    matrix: str
    locations: str
    knots: str
    widths: str


@dataclass
class ReversibleJump(MetropolisHastings):
    """ref: reversible_jump.py:24-74 (the whole model stays attached: no conditional model is taken)."""

    associated_params: Union[list, str, None] = None
    n_max: Union[int, None] = None
    birth_probability: float = 0.5
    state_birth_function: Union[Callable, None] = None
    state_death_function: Union[Callable, None] = None
    matching_params: Union[dict, None] = None
    basis: Union[GaussianKernelBasis, None] = None

    def __post_init__(self):
        if isinstance(self.associated_params, str):
            self.associated_params = [self.associated_params]

    # names of the state entries this sampler rewrites besides self.param (downloaded / stored with it)
    def extra_state_names(self):
        b = self.basis
        return [b.knots, b.widths, self.matching_params["variable"], b.matrix] if b is not None else []

    def stored_state_names(self):
        """Entries stored per iteration next to self.param (the basis matrix is state, not a sample)."""
        b = self.basis
        return [b.knots, b.widths, self.matching_params["variable"]] if b is not None else []

    def trim_host_state(self, state: dict, fetch):
        """Padded device state (one chain) -> the reference's exact shapes in `state`: count scalar, knots / widths
        (1, n), coefficients (n, 1), basis (n_data, n).  The allocation vector of a mixture prior on the coefficients
        follows the count, as the reference's birth / death callbacks grow / shrink it
        (tests/test_reversible_jump.py:91,117 of the reference).  `fetch(name)` returns the padded host copy."""
        from openmcmc_b200.parameter import MixtureParameterVector

        b, var = self.basis, self.matching_params["variable"]
        n = int(np.ravel(np.asarray(fetch(self.param)))[0])
        state[b.knots] = np.asarray(fetch(b.knots)).reshape(1, -1)[:, :n]
        state[b.widths] = np.asarray(fetch(b.widths)).reshape(1, -1)[:, :n]
        state[var] = np.asarray(fetch(var)).reshape(-1, 1)[:n]
        Bm = np.asarray(fetch(b.matrix))
        state[b.matrix] = Bm.reshape(-1, Bm.shape[-1])[:, :n]
        prior = self.model.get(var) if hasattr(self.model, "get") else None
        if prior is not None and isinstance(getattr(prior, "mean", None), MixtureParameterVector):
            alloc = prior.mean.allocation
            if alloc in state and np.size(state[alloc]) != n:
                state[alloc] = np.zeros((n, 1), dtype=np.asarray(state[alloc]).dtype)
        return n

    def _pattern(self, plan, host_state):
        """Match the attached model onto the kernel's fixed structure; PlanError for anything else."""
        from openmcmc_b200.distribution.distribution import Gamma, Poisson, Uniform
        from openmcmc_b200.distribution.location_scale import Normal, NullDistribution
        from openmcmc_b200.parameter import Identity, LinearCombination, MixtureParameterMatrix, MixtureParameterVector, ScaledMatrix

        if self.basis is None:
            raise engine.PlanError(
                "ReversibleJump on the device needs basis=GaussianKernelBasis(matrix, locations, knots, widths); Python "
                "state_birth_function / state_death_function callbacks cannot run inside a kernel (SURVEY F10)")
        if self.matching_params is None:
            raise engine.PlanError("ReversibleJump on the device needs matching_params (matched coefficient transitions)")
        if self.n_max is None or int(self.n_max) < 2:
            raise engine.PlanError("ReversibleJump needs n_max >= 2")
        b, mp = self.basis, self.matching_params
        var, mat = mp["variable"], mp["matrix"]
        if mat != b.matrix:
            raise engine.PlanError("matching_params['matrix'] must be the basis matrix")
        assoc = list(self.associated_params or [])
        if b.knots not in assoc or any(a not in (b.knots, b.widths) for a in assoc):
            raise engine.PlanError("associated_params must be the basis knots (and optionally the widths)")
        out = dict(var=var, sample_omega=b.widths in assoc, y=None, tau_y=None)
        seen = set()
        for name, d in self.model.items():
            if isinstance(d, (Normal, NullDistribution)) and isinstance(d.mean, LinearCombination) and d.mean.form == {var: mat}:
                if type(d) is Normal:
                    if not isinstance(d.precision, ScaledMatrix):
                        raise engine.PlanError("RJ response precision must be ScaledMatrix(identity, scalar)")
                    P = engine.ensure_matrix(plan.state, host_state, d.precision.matrix)
                    if P.kind != "eye":
                        raise engine.PlanError("RJ response precision matrix must be the identity")
                    out.update(y=d.response, tau_y=d.precision.scalar)
                seen.add("response")
            elif isinstance(d, Normal) and d.response == var:
                mean, prec = d.mean, d.precision
                if not isinstance(mean, (MixtureParameterVector, Identity)) or not isinstance(prec, (MixtureParameterMatrix, Identity)):
                    raise engine.PlanError("RJ coefficient prior must be Normal with (mixture) scalar mean and precision")
                mname = mean.param if isinstance(mean, MixtureParameterVector) else mean.form
                pname = prec.param if isinstance(prec, MixtureParameterMatrix) else prec.form
                if np.size(host_state[mname]) != 1 or np.size(host_state[pname]) != 1:
                    raise engine.PlanError("RJ coefficient prior: one mixture component (scalar mean / precision) only")
                out.update(mu_beta=mname, tau_beta=pname)
                seen.add("prior")
            elif isinstance(d, Poisson) and d.response == self.param:
                out["rho"] = d.rate.form
                seen.add("count")
            elif isinstance(d, Uniform) and d.response == b.knots:
                out.update(theta_lo=float(np.ravel(d.domain_response_lower)[0]), theta_hi=float(np.ravel(d.domain_response_upper)[0]))
                seen.add("knots")
            elif isinstance(d, Gamma) and d.response == b.widths:
                out.update(omega_shape=d.shape.form, omega_rate=d.rate.form)
                seen.add("widths")
            else:
                raise engine.PlanError(f"ReversibleJump: distribution of '{name}' is outside the device RJ model")
        need = {"response", "prior", "count", "knots"} | ({"widths"} if out["sample_omega"] else set())
        if not need <= seen:
            raise engine.PlanError(f"ReversibleJump: model lacks {sorted(need - seen)}")
        return out

    def setup(self, plan, host_state, debug_draws=None):
        """Create the padded device state and the omc_rj_t description once per plan.  Called by MCMC.prepare before
        any sampler compiles, so that the companion samplers of the RJ model (ManifoldMALA on the coefficients,
        RandomWalkLoop on knots / widths) find the padded state whatever the sampler order."""
        import torch

        st = plan.state
        C, dev = st.n_chains, st.device
        ctx = plan.ctx(self)
        if "args" not in ctx:
            pt = self._pattern(plan, host_state)
            b, n_max = self.basis, int(self.n_max)
            X = st[b.locations]
            nd = X.size
            n0 = np.asarray(host_state[self.param], dtype=np.float64).reshape(-1)
            n_dev = torch.as_tensor(np.broadcast_to(n0, (C,)).copy() if n0.size == 1 else n0).to(dev)

            def padded(name, rows_fill):
                a = np.asarray(host_state[name], dtype=np.float64)
                a = a.reshape(C, -1) if a.ndim == 3 else np.broadcast_to(a.reshape(1, -1), (C, a.size))
                out = np.full((C, n_max), rows_fill)
                out[:, : a.shape[1]] = a
                return torch.as_tensor(out).to(dev)

            theta, omega, beta = padded(b.knots, 0.0), padded(b.widths, 1.0), padded(pt["var"], 0.0)
            Bm = torch.zeros(C, nd, n_max, dtype=torch.float64, device=dev)
            for name, t, rows, cols in ((self.param, n_dev, 1, 1), (b.knots, theta, 1, n_max), (b.widths, omega, 1, n_max),
                                        (pt["var"], beta, n_max, 1), (b.matrix, Bm, nd, n_max)):
                if name in st.arrays:
                    st._retired.append(st.arrays[name])
                st.arrays[name] = engine.DevArray(t.reshape(C, rows, cols), True, rows, cols)
            ctx["rng"] = plan.rng_site()
            ctx["counters"] = torch.zeros(C, 2, dtype=torch.int64, device=dev)
            ctx["probe"] = plan.new(C, 8, fill=0.0) if plan.probes is not None and plan.probes.get("enable") else None
            if ctx["probe"] is not None:
                plan.probes[self.param] = {"step": ctx["probe"]}
            dbg, dbg_stride = None, 0
            if debug_draws and "rj" in debug_draws:
                dbg, dbg_stride = plan.debug_tensor(debug_draws["rj"], 6)
            lim = self.matching_params.get("limits")
            v = lambda key: st[pt[key]].vec() if pt.get(key) else None  # noqa: E731
            yv = None
            if pt["y"] is not None:
                ya = st[pt["y"]]
                yv = ya.vec()
            live = LIVE_GRAM and n_max <= 128
            ctx["gram_valid"] = plan.keep_tensor(torch.zeros(C, dtype=torch.int32, device=dev)) if live else None
            ctx["args"] = K.rj_args(
                C, nd, n_max, n_dev, theta, omega, beta, Bm, X.data, pt["theta_lo"], pt["theta_hi"],
                self.birth_probability, y=yv, tau_y=v("tau_y"), sample_omega=pt["sample_omega"],
                omega_shape=v("omega_shape"), omega_rate=v("omega_rate"), mu_beta=v("mu_beta"), tau_beta=v("tau_beta"),
                rho=v("rho"), match_scale=float(self.matching_params["scale"]), match_limits=lim, rng_=ctx["rng"],
                debug=dbg, debug_sweep_stride=dbg_stride, counters=ctx["counters"], status=plan.status,
                probe=ctx["probe"], size_class=plan.keep_tensor(torch.zeros(C, dtype=torch.int32, device=dev)),
                # the live Gram matrix B'B in the chain state (omc.h): a birth costs one new column of inner products
                gram=plan.keep_tensor(torch.empty(C, n_max, n_max, dtype=torch.float64, device=dev)) if live else None,
                gram_valid=ctx["gram_valid"])
            plan.keep.extend([n_dev, theta, omega, beta, Bm])
            K.rj_basis(ctx["args"])     # basis of the initial knots (the host copy is not trusted to be padded)
            plan.__dict__["_rj"] = self
        return ctx["args"]

    def reset_derived(self, plan):
        """After the chain state was put back by hand (the eager warm-up pass of MCMC.prepare): the live Gram matrices
        belong to the state that was there before."""
        gv = plan.ctx(self).get("gram_valid")
        if gv is not None:
            gv.zero_()

    def compile(self, plan, host_state, debug_draws=None):
        args = self.setup(plan, host_state, debug_draws)
        plan.emit(lambda: K.reversible_jump(args), f"reversible_jump[{self.param}]")
        for name in [self.param] + self.extra_state_names():
            plan.wrote(name)

    def compile_log_post(self, plan, out):
        """model.log_p(state) for the RJ model, one value per chain (mcmc.py:108)."""
        args = plan.ctx(self)["args"]

        def launch():
            args.logp_out = out.data_ptr()
            K.reversible_jump(args, logp_only=True)

        plan.emit(launch, f"rj_log_post[{self.param}]")
