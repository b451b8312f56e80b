"""MCMC driver (host-side mirror of `openmcmc.mcmc.MCMC`).  ref: mcmc.py:18-115

Same constructor and `run_mcmc()` / `.store` / `.state` surface as the reference, plus keyword extras for the batched
device engine: `n_chains`, `seed`, `device`, `chain_offset`, `debug_draws`.  `run_mcmc()` uploads the state once,
compiles the (model, samplers) pair into a sweep plan, captures one sweep as a CUDA graph, replays it
(n_burn + n_iter) * n_thin times with the chain state resident in HBM, and brings the stored samples back.
"""

from copy import copy
from dataclasses import dataclass, field

import numpy as np
import torch
from scipy import sparse

from openmcmc_b200 import engine
from openmcmc_b200 import kernels as K
from openmcmc_b200.model import Model


class LazyHostArray:
    """A final-state entry that stays on the device until somebody looks at it (`np.asarray`, indexing, `.shape`).

    Used for the padded basis matrix of a multi-chain ReversibleJump run ([C, n_data, n_max]: 4.3 GB at the C5 size, a
    quantity derived from knots and widths that most callers never read)."""

    def __init__(self, fetch, shape):
        self._fetch, self._value, self.shape = fetch, None, tuple(shape)
        self.ndim, self.dtype = len(self.shape), np.dtype(np.float64)

    def __array__(self, dtype=None, copy=None):
        if self._value is None:
            self._value = self._fetch()
        return self._value if dtype is None else self._value.astype(dtype)

    def __getitem__(self, idx):
        return self.__array__()[idx]

    def __len__(self):
        return self.shape[0]


STREAM_STORE_BYTES = 256 << 20  # device stores above this are streamed to the host during the run (stream_store=None)
UPLOAD_PIECE_BYTES = 256 << 20  # chain blocks are uploaded in pieces of this size (see _run_blocked.stage)
FUSE_STORED_SWEEP = True       # n_thin = 1: sweep + store epilogue captured as one graph, small ops fused across
RING_BYTES = 1 << 30           # budget of the device ring of a streamed store (at least 2 slabs; the drain, not the ring, is the limit)


@dataclass
class MCMC:
    """ref: mcmc.py:18-85.  For n_chains == 1 `store[param]` has the reference shape (size, n_iter); for n_chains > 1
    it is (n_chains, size, n_iter) and `store["log_post"]` is (n_iter, n_chains)."""

    state: dict
    samplers: list
    model: Model
    n_burn: int = 5000
    n_iter: int = 5000
    n_thin: int = 1
    n_chains: int = 1
    seed: int = 0
    device: int = None
    chain_offset: int = 0
    debug_draws: dict = None   # {param: {"z"|"g"|"u": array [n_sweeps, (C,) size]}} injected random streams
    probes: bool = False       # keep per-sampler intermediates (Q, b, L, mu, a*, b*) of the LAST sweep
    stream_store: bool = None  # True: stored iterations leave the device as they are produced (device ring -> pinned
    #                            staging -> host arrays on a copy stream, stream_store.py) and n_iter is no longer
    #                            bounded by HBM; False: resident device store, one download at the end; None: stream when
    #                            the store would exceed STREAM_STORE_BYTES
    upload_blocks: int = None  # > 1: run the chains as that many contiguous chain blocks, block k+1 being uploaded and
    #                            compiled while block k sweeps (chains are independent and the RNG is keyed by the
    #                            global chain id, so the draws are the same); None = one block per 2.7 GB of per-chain
    #                            HOST input, at most 16; see _run_blocked
    store: dict = field(default_factory=dict, init=False)

    def __post_init__(self):
        """State coercion as mcmc.py:63-76 (host side: shapes only).  Missing sampled parameters are drawn from their
        prior on the device (mcmc.py:78-80)."""
        self.state = copy(self.state)
        for key, term in self.state.items():
            if sparse.issparse(term) or isinstance(term, torch.Tensor):
                continue
            if not isinstance(term, np.ndarray):
                term = np.array(term, ndmin=2, dtype=np.float64)
                if np.shape(term)[0] == 1:
                    term = term.T
            elif term.ndim < 2:
                term = np.atleast_2d(term).T
            self.state[key] = term
        # Start values missing from the state are prior draws (mcmc.py:78-80) -- one PER CHAIN, column = global chain id,
        # so that chains start over-dispersed (what split-R-hat assumes) and a sharded run starts where the unsharded one
        # does.  The host state keeps the first local chain's draw (shapes, later hierarchical draws); the per-chain
        # values go to the device in _prepare.
        self._chain_starts = {}
        for sampler in self.samplers:
            if sampler.param not in self.state:
                dist = sampler.model[sampler.param]
                if self.n_chains > 1:
                    draws = np.asarray(dist.rvs(self.state, n=self.chain_offset + self.n_chains), dtype=np.float64)
                    draws = draws[:, self.chain_offset:]
                    self.state[sampler.param] = draws[:, :1].copy()
                    self._chain_starts[sampler.param] = np.ascontiguousarray(draws.T)[:, :, None]
                else:
                    self.state[sampler.param] = dist.rvs(self.state)
        self._prepared = None
        self.status = None
        self.timing = {}
        self._init_store()

    def _init_store(self):
        """ref: mcmc.py:81-85 -- NaN-filled sample arrays exist before the run (through sampler.init_store, reference
        shapes; a leading chain axis when n_chains > 1).  Skipped above 64 MB: the run allocates what it fills."""
        total = 0
        for s in self.samplers:
            v = self.state.get(s.param)
            total += int(np.size(v)) if v is not None and not isinstance(v, torch.Tensor) else 0
        if total == 0 or total * max(self.n_iter, 1) * self.n_chains * 8 > (64 << 20):
            return
        try:
            one = {}
            for s in self.samplers:
                one = s.init_store(current_state=self.state, store=one, n_iterations=self.n_iter)
            one["log_post"] = np.full((self.n_iter, 1), np.nan)
            if self.model.response is not None:
                for response in self.model.response:
                    one[response] = np.full((np.size(self.state[response]), self.n_iter), np.nan)
        except Exception:      # a sampler without host-side shapes yet (e.g. per-chain tensors): the run fills the store
            return
        if self.n_chains == 1:
            self.store = one
        else:
            C = self.n_chains
            self.store = {k: (np.full((self.n_iter, C), np.nan) if k == "log_post" else np.broadcast_to(
                v, (C,) + v.shape).copy()) for k, v in one.items()}

    # ------------------------------------------------------------------ device pipeline
    def prepare(self, warm_up=True, wait=True):
        """Upload + compile + capture.  Returns self (idempotent).  `warm_up=False` skips the eager pass in front of
        the captures (for a plan whose kernels an earlier plan of this process has launched already: the chain blocks
        after the first), `wait=False` leaves the prologue running on the engine's stream."""
        if self._prepared is not None:
            return self
        dev = K.init_device(self.device)
        C = self.n_chains
        sampled = {s.param for s in self.samplers}
        # replicated data responses (y of shape (dim, n_rep)) are compiled in their single-column form
        # (engine.unreplicate); the user's model, samplers and state come back unchanged after the run
        user_model, user_sampler_models = self.model, [s.model for s in self.samplers]
        dists, state_c, changed = engine.unreplicate(list(self.model.values()), self.state, frozenset(sampled))
        if changed:
            twin = {id(a): b for a, b in zip(self.model.values(), dists)}
            self._rep_restore = {k: self.state.get(k) for k in state_c if state_c[k] is not self.state.get(k)}
            self.state = state_c
            self.model = Model(dists, response=self.model.response)
            for s in self.samplers:
                s.model = Model([twin.get(id(d), d) for d in s.model.values()], response=getattr(s.model, "response", None))
        try:
            return self._prepare(dev, C, sampled, warm_up, wait)
        finally:
            self.model = user_model
            for s, m in zip(self.samplers, user_sampler_models):
                s.model = m

    def _prepare(self, dev, C, sampled, warm_up, wait):
        st = engine.DeviceState(C, dev, self.state, per_chain_names=sampled)
        for name, arr in getattr(self, "_chain_starts", {}).items():   # per-chain prior draws of missing start values
            st.put(name, arr, per_chain=True)
        self.stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self.stream):
            plan = engine.Plan(st, seed=self.seed, chain_offset=self.chain_offset)
            plan.probes = {"enable": True} if self.probes else {}
            dd = self.debug_draws or {}
            from openmcmc_b200 import gmrf_plan

            gmrf_plan.discover(plan, self.state, list(self.model.values()))
            rj = self._rj_sampler()
            if rj is not None:
                rj.setup(plan, self.state, dd.get(rj.param))
            # pass A (dry): learn which derived quantities are valid at the end of a sweep
            plan.ops = []
            for s in self.samplers:
                s.compile(plan, self.state, dd.get(s.param))
            valid_end = dict(plan.valid)
            # pass B: steady-state sweep, assuming the end-of-sweep validity at its start
            plan.ops = sweep_ops = []
            for s in self.samplers:
                s.compile(plan, self.state, dd.get(s.param))
            if plan.valid != valid_end:  # no fixed point (should not happen): fall back to the cold sweep
                plan.valid = {k: False for k in plan.valid}
                plan.ops = sweep_ops = []
                for s in self.samplers:
                    s.compile(plan, self.state, dd.get(s.param))
                valid_end = {k: False for k in plan.valid}
            sweep_ops.append(("sweep_counter", lambda: K.counter_add(plan.sweep_counter, 1)))
            # store epilogue (ref: mcmc.py:105-111): resident [n_iter, C, size] buffers, or -- streamed -- a ring of a few
            # slabs that stream_store.StoreStreamer drains to the host while the next sweeps run
            plan.ops = store_ops = []
            n_iter = max(self.n_iter, 1)
            self._dev_store = {}
            self._store_names = []
            stored = []
            for s in self.samplers:
                for name in [s.param] + list(getattr(s, "stored_state_names", lambda: [])()):
                    if name not in [nm for nm, _ in stored]:
                        stored.append((name, st[name]))
            slab = 8 * C * (sum(arr.size for _, arr in stored) + 1)
            if self.model.response is not None:
                slab += 8 * C * sum(st[r].rows for r in self.model.response)
            streamed = self.stream_store if self.stream_store is not None else slab * n_iter > STREAM_STORE_BYTES
            self._streamed = bool(streamed) and self.n_iter >= 1
            self._ring = max(2, min(8, int(RING_BYTES // max(slab, 1)))) if self._streamed else 0
            self._slab_bytes = slab
            rows = self._ring if self._streamed else n_iter
            ring = self._streamed
            if self._streamed:
                from openmcmc_b200 import stream_store

                self._staging_warm = stream_store.warm(dev)
            for name, arr in stored:
                buf = plan.new(rows, C, arr.size, fill=float("nan"))
                self._dev_store[name] = buf
                self._store_names.append(name)
                plan.emit((lambda arr=arr, buf=buf: K.store_copy(arr.data, buf, C * arr.size, plan.iter_counter, rows, ring=ring)),
                          f"store[{name}]",
                          fop=(lambda arr=arr, buf=buf: K.fop_store_copy(arr.data, buf, C * arr.size, plan.iter_counter, rows,
                                                                         ring=ring)) if C * arr.size <= K.FUSED_COPY_MAX else None)
            self._dev_logpost = plan.new(rows, C, fill=float("nan"))
            self._logpost_now = plan.new(C)
            saved_valid = dict(plan.valid)
            rj = plan.__dict__.get("_rj")
            if rj is not None:
                rj.compile_log_post(plan, self._logpost_now)
            else:
                engine.compile_log_post(plan, self.state, self.model, self._logpost_now)
            plan.emit((lambda: K.store_copy(self._logpost_now, self._dev_logpost, C, plan.iter_counter, rows, ring=ring)),
                      "store[log_post]",
                      fop=(lambda: K.fop_store_copy(self._logpost_now, self._dev_logpost, C, plan.iter_counter, rows, ring=ring)))
            self._dev_fitted = {}
            if self.model.response is not None:
                for response, predictor in self.model.response.items():
                    self._dev_fitted[response] = engine.compile_fitted(plan, self.state, self.model[response],
                                                                       predictor, rows, ring=ring)
            store_ops.append(("iter_counter", lambda: K.counter_add(plan.iter_counter, 1)))
            plan.valid = saved_valid
            # prologue: quantities the steady-state sweep assumes valid, computed from the initial state
            plan.ops = prologue_ops = []
            plan.valid = {k: False for k in plan.valid}    # nothing is valid yet: every compute() emits its full form
            for qname, ok in valid_end.items():
                if ok:
                    plan.require(qname)                    # (a compute() may require other quantities first)
            plan.valid = saved_valid
            self.plan = plan
            counter = sweep_ops.pop()                           # the sweep counter stays a launch of its own, last
            sweep_raw = list(sweep_ops)
            sweep_ops = engine.fuse_small_ops(plan, sweep_ops) + [counter]
            counter = store_ops.pop()
            store_raw = list(store_ops)
            store_ops = engine.fuse_small_ops(plan, store_ops) + [counter]
            self._ops = {"prologue": prologue_ops, "sweep": sweep_ops, "store": store_ops}
            # n_thin = 1: every sweep is a stored one, so sweep + store epilogue are ONE graph whose small per-chain ops
            # (conjugate scalar draws, log-density terms, store copies) fuse across the boundary, both counters last
            stored_sweep_ops = None
            if FUSE_STORED_SWEEP and self.n_thin == 1 and not self._streamed and self.n_iter >= 1:
                stored_sweep_ops = engine.fuse_small_ops(plan, sweep_raw + store_raw) + [
                    ("counters", lambda: K.counter_add2(plan.sweep_counter, 1, plan.iter_counter, 1))]
                self._ops["stored_sweep"] = stored_sweep_ops
            if warm_up:
                self._warm_up(plan, st, sampled)
            self._sweep_graph = K.Graph.capture(lambda: [fn() for _, fn in sweep_ops])
            self._store_graph = K.Graph.capture(lambda: [fn() for _, fn in store_ops])
            self._stored_sweep_graph = (K.Graph.capture(lambda: [fn() for _, fn in stored_sweep_ops])
                                        if stored_sweep_ops is not None else None)
            for _, fn in prologue_ops:
                fn()
        if wait:
            self.stream.synchronize()
        self._prepared = st
        return self

    def _warm_up(self, plan, st, sampled):
        """Run every op of the prologue, the sweep and the store epilogue once, eagerly, then put the chain state back.

        A kernel's first launch can make the driver load its module and grow the context's local-memory pool; both are
        synchronising operations that invalidate a stream capture, so they have to happen before the graphs are
        recorded.  Restored afterwards: sampled parameters, sweep / iteration counters, status bits, MH accept counters
        (derived quantities are recomputed by the prologue, store rows are overwritten by the first stored iteration).
        """
        names = set(sampled)
        for s in self.samplers:
            names.update(getattr(s, "extra_state_names", lambda: [])())
        saved = [(st[name].data, st[name].data.clone()) for name in names]
        for ctx in plan.__dict__.get("_ctx", {}).values():
            if ctx.get("counters") is not None:
                saved.append((ctx["counters"], ctx["counters"].clone()))
        for t in (plan.sweep_counter, plan.iter_counter, plan.status):
            saved.append((t, t.clone()))
        for phase in ("prologue", "sweep", "store", "stored_sweep"):
            for _, fn in self._ops.get(phase, []):
                fn()
        for t, copy_ in saved:
            t.copy_(copy_)
        for s in self.samplers:       # derived chain state a sampler keeps besides the state entries (live Gram matrix)
            hook = getattr(s, "reset_derived", None)
            if hook is not None:
                hook(plan)
        torch.cuda.current_stream().synchronize()

    def launches_per_sweep(self) -> int:
        self.prepare()
        return self._sweep_graph.num_kernels()

    def launches_of(self, n_burn: int, n_iter: int, n_thin: int) -> int:
        """Kernel launches run_device(n_burn, n_iter, n_thin) replays."""
        self.prepare()
        sw, store = self._sweep_graph.num_kernels(), self._store_graph.num_kernels()
        if n_thin == 1 and n_iter > 0 and not self._streamed and self._stored_sweep_graph is not None:
            return n_burn * sw + n_iter * self._stored_sweep_graph.num_kernels()
        return (n_burn + n_iter * n_thin) * sw + n_iter * store

    def run_device(self, n_burn=None, n_iter=None, n_thin=None, restart_store=True):
        """Replay the captured sweep graph on the engine's stream (asynchronous).  restart_store=False continues the
        store where the previous call left it (a run split into several calls)."""
        self.prepare()
        n_burn = self.n_burn if n_burn is None else n_burn
        n_iter = self.n_iter if n_iter is None else n_iter
        n_thin = self.n_thin if n_thin is None else n_thin
        with torch.cuda.stream(self.stream):
            if restart_store:
                self.plan.iter_counter.zero_() # a run stores from iteration 0 again, as the reference overwrites its store
            if self._streamed and n_iter > 0:
                return self._run_streamed(n_burn, n_iter, n_thin)
            if n_thin == 1 and n_iter > 0 and getattr(self, "_stored_sweep_graph", None) is not None:
                if n_burn:
                    K.run_schedule(self._sweep_graph, None, n_burn, 0, 1)
                self._stored_sweep_graph.launch(n_iter)
                return
            K.run_schedule(self._sweep_graph, self._store_graph, n_burn, n_iter, n_thin)

    def _run_streamed(self, n_burn, n_iter, n_thin):
        """The schedule of omc_run_schedule with the store graph of every stored iteration fenced against the copy
        stream that drains the device ring (stream_store.py).  Returns when the last slab has landed in host memory."""
        from openmcmc_b200 import stream_store

        st, C = self._prepared, self.n_chains
        self._staging_warm.join()
        self._host_store = {name: stream_store.host_array((n_iter, C, st[name].size)) for name in self._store_names}
        self._host_logpost = np.empty((n_iter, C))
        self._host_fitted = {r: stream_store.host_array((n_iter, C, buf.shape[-1])) for r, buf in self._dev_fitted.items()}
        entries = [(self._dev_store[name], self._host_store[name]) for name in self._store_names]
        entries.append((self._dev_logpost, self._host_logpost))
        entries += [(self._dev_fitted[r], self._host_fitted[r]) for r in self._dev_fitted]
        streamer = stream_store.StoreStreamer(st.device, entries, self._ring, n_iter)
        try:
            if n_burn:
                K.run_schedule(self._sweep_graph, None, n_burn, 0, n_thin)
            for it in range(n_iter):
                K.run_schedule(self._sweep_graph, None, 1, 0, n_thin)     # the n_thin sweeps of stored iteration `it`
                streamer.before_store(it, self.stream)
                self._store_graph.launch(1)
                streamer.after_store(it, self.stream)
        finally:
            streamer.finish()
        self._streamed_rows = n_iter
        self.timing["streamed_d2h_bytes"] = streamer.d2h_bytes

    def collect(self):
        """Download the stored samples and the final state (ref shapes; see class docstring)."""
        st = self._prepared
        self.stream.synchronize()
        C = self.n_chains
        n_done = int(self.plan.iter_counter.item())
        self.store = {}
        d2h = 0
        owner = {}
        for s in self.samplers:
            for name in [s.param] + list(getattr(s, "extra_state_names", lambda: [])()):
                owner.setdefault(name, s)
        streamed = self._streamed and getattr(self, "_streamed_rows", 0) > 0
        last_rows = {}
        if streamed:
            self._mask_padded_store_host()
        else:
            self._mask_padded_store_device()
        for name in self._store_names:
            s = owner[name]
            buf = self._host_store[name] if streamed else K.download(self._dev_store[name][: self.n_iter])  # [n_iter, C, size]
            d2h += buf.nbytes
            if self.n_iter > 0:
                last_rows[name] = buf[self.n_iter - 1]
            arr = np.transpose(buf, (1, 2, 0))                                 # [C, size, n_iter]
            if name == s.param:
                arr = self._shape_store(s, arr, st[name])
            self.store[name] = arr[0] if C == 1 else arr
        lp = self._host_logpost if streamed else K.download(self._dev_logpost[: self.n_iter])
        d2h += lp.nbytes
        self.store["log_post"] = lp.reshape(self.n_iter, 1) if C == 1 else lp
        for response, buf in self._dev_fitted.items():
            h = self._host_fitted[response] if streamed else K.download(buf[: self.n_iter])
            d2h += h.nbytes
            arr = np.transpose(h, (1, 2, 0))
            self.store[response] = arr[0] if C == 1 else arr
        rj = self._rj_sampler()
        lazy = {rj.basis.matrix} if (rj is not None and rj.basis is not None and C > 1) else set()
        for s in self.samplers:
            for name in [s.param] + list(getattr(s, "extra_state_names", lambda: [])()):
                if name in lazy:     # derived from knots / widths: downloaded when somebody reads it
                    arr = st.arrays[name]
                    self.state[name] = LazyHostArray((lambda name=name: st.get_host(name)), arr.data.shape)
                    continue
                if name in last_rows and rj is None and n_done == self.n_iter:
                    # the state after the last sweep IS the last stored iteration: no second download (C3: 512 MB)
                    a = st.arrays[name]
                    # (a view of the store's last row: a copy of a 512 MB field would cost what the download did)
                    new = last_rows[name].reshape((C, a.rows, a.cols) if C > 1 else (a.rows, a.cols))
                else:
                    new = st.get_host(name)
                    d2h += new.nbytes
                self.state[name] = new
        self._trim_padded_state()
        for key, original in getattr(self, "_rep_restore", {}).items():   # the user's replicated arrays come back
            if original is None:
                self.state.pop(key, None)
            else:
                self.state[key] = original
        self.status = self.plan.status.cpu().numpy()
        self.timing["d2h_bytes"] = d2h
        self.timing["h2d_bytes"] = st.h2d_bytes
        self.timing["stored_iterations"] = n_done
        return self.store

    def _rj_sampler(self):
        from openmcmc_b200.sampler.reversible_jump import ReversibleJump

        rj = next((s for s in self.samplers if isinstance(s, ReversibleJump)), None)
        if rj is None:   # a companion sampler of an RJ model run on its own names its RJ sampler explicitly
            rj = next((s.rj for s in self.samplers if getattr(s, "rj", None) is not None), None)
        return rj

    def _mask_padded_store_device(self):
        """Variable-dimension parameters are stored at capacity n_max: entries beyond the stored count become NaN, the
        layout `max_variable_size` gives the reference's store (sampler.py:69-118).  Done on the device store
        [n_iter, C, size] before the download (on the host it was a boolean-mask assignment over 6e7 elements per
        parameter, most of the C5 collect time)."""
        rj = self._rj_sampler()
        if rj is None or rj.param not in self._dev_store:
            return
        with torch.cuda.stream(self.stream):
            cnt = self._dev_store[rj.param][: self.n_iter]                  # [n_iter, C, 1]
            for name in rj.stored_state_names():
                if name in self._dev_store:
                    buf = self._dev_store[name][: self.n_iter]
                    idx = torch.arange(buf.shape[-1], device=buf.device, dtype=buf.dtype).view(1, 1, -1)
                    buf.masked_fill_(idx >= cnt, float("nan"))
        self.stream.synchronize()

    def _mask_padded_store_host(self):
        """The same NaN mask on a streamed store (host arrays [n_iter, C, size])."""
        rj = self._rj_sampler()
        if rj is None or rj.param not in self._host_store:
            return
        cnt = self._host_store[rj.param]                                    # [n_iter, C, 1]
        for name in rj.stored_state_names():
            if name in self._host_store:
                buf = self._host_store[name]
                buf[np.arange(buf.shape[-1]).reshape(1, 1, -1) >= cnt] = np.nan

    def device_samples(self, name, elem_stride=1):
        """Stored draws of `name` as a device tensor [n_stored, C, n_sel] for the diagnostics kernels: a view of the
        resident store, or -- streamed -- the strided selection of the host store uploaded again."""
        n_done = int(self.plan.iter_counter.item())
        if self._streamed and getattr(self, "_streamed_rows", 0) > 0:
            sel = np.ascontiguousarray(self._host_store[name][:n_done, :, ::elem_stride])
            return K.upload(sel, self._prepared.device), 1
        return self._dev_store[name][:n_done], elem_stride

    def _trim_padded_state(self):
        """Final state of one chain in the reference's exact shapes (theta (1,n), beta (n,1), B (n_data,n))."""
        rj = self._rj_sampler()
        if rj is None or self.n_chains != 1:
            return
        rj.trim_host_state(self.state, lambda name: self.state[name])

    @staticmethod
    def _shape_store(sampler, arr, dev_arr):
        """max_variable_size layouts of the reference's sample store (sampler.py:69-118): tuple -> the parameter's own
        (rows, cols) shape padded with NaN; int -> flat, padded with NaN.  arr: [C, size, n_iter]."""
        mvs = sampler.max_variable_size
        if mvs is None:
            return arr
        C, size, n_iter = arr.shape
        if isinstance(mvs, tuple):
            out = np.full((C,) + tuple(mvs) + (n_iter,), np.nan)
            out[:, : dev_arr.rows, : dev_arr.cols, :] = arr.reshape(C, dev_arr.rows, dev_arr.cols, n_iter)
            return out
        if size == int(mvs):      # stored at capacity already (padded ReversibleJump state): nothing to pad
            return arr
        out = np.full((C, int(mvs), n_iter), np.nan)
        out[:, :size, :] = arr
        return out

    # ------------------------------------------------------------------ chain blocks: upload under compute
    def _n_blocks(self):
        """Chain blocks run_mcmc() uses: `upload_blocks`, limited so that every block keeps at least 2 chains; 1 when
        the run injects draws / keeps probes / holds a ReversibleJump sampler (their host-side set-up is per run)."""
        B = self.upload_blocks
        if B is None:   # automatic: only per-chain inputs that still live on the host count
            C, host_bytes = self.n_chains, 0
            for v in self.state.values():
                if isinstance(v, np.ndarray) and v.ndim == 3 and v.shape[0] == C:
                    host_bytes += v.nbytes
                elif isinstance(v, torch.Tensor) and v.dim() == 3 and v.shape[0] == C and not v.is_cuda:
                    host_bytes += v.numel() * v.element_size()
            B = min(16, int(host_bytes // 2.7e9))   # ~60 ms of PCIe per block: twice the host's per-block plan / capture time
        B = min(int(B), self.n_chains // 2)
        if B <= 1 or self.debug_draws or self.probes or self._rj_sampler() is not None:
            return 1
        return B

    def _run_blocked(self, B):
        """The chains as B contiguous blocks, each a sub-run with `chain_offset` = its first global chain id.  A block's
        host->device upload, plan compile and graph capture happen on the host thread while the previous blocks' sweeps
        run asynchronously on their own streams, so the PCIe transfer of X hides under the sweeps (and vice versa).
        Results are merged into the shapes of an unblocked run and equal it bit for bit as long as the blocks pick the
        same row split in omc_reg_pass (always the case once a block has a few hundred chains)."""
        import time

        C = self.n_chains
        bounds = [(C * b // B, C * (b + 1) // B) for b in range(B)]
        t0 = time.perf_counter()
        dev = torch.device("cuda", K.init_device(self.device))
        copy_stream = torch.cuda.Stream(device=dev)

        def per_chain(v):
            return isinstance(v, (np.ndarray, torch.Tensor)) and v.ndim == 3 and v.shape[0] == C

        arenas = {}

        def stage(lo, hi):
            """Queue the host->device copies of a block's pinned per-chain tensors on the copy stream.  The copies of
            block k+1 are queued BEFORE the host turns to block k's plan, so the PCIe link never waits for Python."""
            out, nbytes = {}, 0
            with torch.cuda.stream(copy_stream):
                for key, v in self.state.items():
                    if per_chain(v) and isinstance(v, torch.Tensor) and not v.is_cuda and v.is_pinned():
                        # in pieces of ~256 MB: the small (pageable) uploads of the block whose plan the host is
                        # compiling meanwhile share the copy engine with this transfer and would otherwise wait behind
                        # all of it (90 ms per block at the C2 shape)
                        if key not in arenas:   # ONE device allocation per input for all blocks (they all stay resident
                            # until the run is collected): the blocks are views, so the caching allocator is asked once
                            # instead of once per block -- a fragmented cache answered 1.4 GB requests with a
                            # synchronising free / malloc cycle of 0.15-0.3 s each
                            arenas[key] = torch.empty((C,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
                        dst = arenas[key][lo:hi]
                        per = max(1, int(UPLOAD_PIECE_BYTES // max(v[0].numel() * v.element_size(), 1)))
                        for i0 in range(lo, hi, per):
                            i1 = min(hi, i0 + per)
                            dst[i0 - lo:i1 - lo].copy_(v[i0:i1], non_blocking=True)
                        out[key] = dst
                        nbytes += dst.numel() * dst.element_size()
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return out, ev, nbytes

        subs, staged_bytes, block_log = [], 0, []
        staged = stage(*bounds[0])
        for b, (lo, hi) in enumerate(bounds):
            on_dev, ev, nbytes = staged
            tb0 = time.perf_counter()
            staged = stage(*bounds[b + 1]) if b + 1 < B else None
            tb1 = time.perf_counter()
            ev.synchronize()       # this block's inputs are in HBM; the next block's are on their way
            tb2 = time.perf_counter()
            staged_bytes += nbytes
            st = {}
            for key, v in self.state.items():
                st[key] = on_dev[key] if key in on_dev else (v[lo:hi] if per_chain(v) else v)
            sub = MCMC(st, self.samplers, self.model, n_burn=self.n_burn, n_iter=self.n_iter, n_thin=self.n_thin,
                       n_chains=hi - lo, seed=self.seed, device=self.device, chain_offset=self.chain_offset + lo,
                       upload_blocks=1, stream_store=False)   # blocks sweep asynchronously: their stores stay resident
            sub._chain_starts = {k: v[lo:hi] for k, v in getattr(self, "_chain_starts", {}).items()}
            # every block keeps the eager warm-up pass in front of its captures: skipping it for the later blocks
            # bought nothing (the host waits behind the next block's transfer anyway) and a capture can be
            # invalidated by the first-time allocations of a block with a different chain count
            sub.prepare()
            sub.run_device()      # asynchronous: the host goes on to the next block
            subs.append(sub)
            block_log.append({"queue_next_s": round(tb1 - tb0, 4), "upload_wait_s": round(tb2 - tb1, 4),
                              "plan_s": round(time.perf_counter() - tb2, 4)})
        t1 = time.perf_counter()
        for sub in subs:
            sub.stream.synchronize()
        t2 = time.perf_counter()
        for sub in subs:
            sub.collect()
        self.store = {}
        for key in subs[0].store:
            axis = 1 if key == "log_post" else 0
            self.store[key] = np.concatenate([sub.store[key] for sub in subs], axis=axis)
        for s in self.samplers:
            for name in [s.param] + list(getattr(s, "extra_state_names", lambda: [])()):
                self.state[name] = np.concatenate([sub.state[name] for sub in subs], axis=0)
        self.status = np.concatenate([sub.status for sub in subs])
        self._blocks = subs
        t3 = time.perf_counter()
        self.timing = {"h2d_bytes": staged_bytes + sum(sub.timing["h2d_bytes"] for sub in subs),
                       "d2h_bytes": sum(sub.timing["d2h_bytes"] for sub in subs),
                       "stored_iterations": subs[0].timing["stored_iterations"], "upload_blocks": B, "blocks": block_log,
                       "prepare_s": t1 - t0, "sweeps_s": t2 - t1, "collect_s": t3 - t2}

    def run_mcmc(self):
        """ref: mcmc.py:87-115"""
        import time

        B = self._n_blocks()
        if B > 1:
            # the host thread is the pacemaker of the block pipeline: a generation-2 pass of Python's cycle collector in
            # the middle of it (0.2-0.3 s in a process that has torch, scipy and a few plans loaded) stalls the uploads
            # queued behind the block being compiled, so collection is paused for the duration of the run
            import gc

            was_enabled = gc.isenabled()
            gc.disable()
            try:
                self._run_blocked(B)
            finally:
                if was_enabled:
                    gc.enable()
            self._finish([sub.plan for sub in self._blocks])
            return
        t0 = time.perf_counter()
        self.prepare()
        t1 = time.perf_counter()
        self.run_device()
        self.stream.synchronize()
        t2 = time.perf_counter()
        self.collect()
        t3 = time.perf_counter()
        self.timing.update({"prepare_s": t1 - t0, "sweeps_s": t2 - t1, "collect_s": t3 - t2})
        self._finish([self.plan])

    def _finish(self, plans):
        """Status and acceptance report of a finished run (ref: mcmc.py:113-115)."""
        if np.any(self.status & 1):
            bad = int(np.sum((self.status & 1) != 0))
            if self.n_chains == 1:
                raise np.linalg.LinAlgError("Matrix is not positive definite")
            print(f"warning: {bad} chain(s) hit a non-positive-definite precision (status bit 1)")
        from openmcmc_b200.sampler.metropolis_hastings import MetropolisHastings

        for sampler in self.samplers:
            if isinstance(sampler, MetropolisHastings):
                counts = []
                for plan in plans:
                    sampler._collect_accept(plan)
                    counts.append(getattr(sampler, "accept_counts", None))
                if len(plans) > 1 and all(c is not None for c in counts):
                    sampler.accept_counts = np.concatenate(counts, axis=0)
                print(f"{sampler.param}: {sampler.accept_rate.get_acceptance_rate()}")
