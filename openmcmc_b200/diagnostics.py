"""Cross-chain diagnostics: per-chain ESS on the device sample store, split-R-hat / total ESS after an all-gather of the
per-chain summaries over the process group (NCCL over NVLink on the GPU box).

The reference has no diagnostics — new functionality, parity unpinned; estimators in csrc/diag.cu, numpy
restatement in oracle/diagnostics.py.  Chains never move between GPUs: each rank summarises its own chains
(`omc_chain_stats`, 64 bytes per chain and scalar parameter), the summaries are all-gathered, and every rank runs
`omc_rhat_combine` on the full set, so all ranks hold identical diagnostics.
"""

import ctypes as C

import torch

from openmcmc_b200 import _cabi
from openmcmc_b200 import kernels as K

RECORD = 8   # n, mean, var, ess, m1, v1, m2, v2


def shard_chains(n_chains_total: int, rank: int, world: int):
    """Contiguous chain block of a rank: (chain_offset, n_local).  Blocks differ by at most one chain."""
    base, rem = divmod(n_chains_total, world)
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def chain_stats(samples: torch.Tensor, elem_stride: int = 1, n_sel: int = None, max_lag: int = 127) -> torch.Tensor:
    """samples: device store [n_iter, n_chains, size] -> [n_chains, n_sel, 8] records (see include/omc.h)."""
    n_iter, n_chains, size = samples.shape
    if n_sel is None:
        n_sel = (size + elem_stride - 1) // elem_stride
    out = torch.empty(n_chains, n_sel, RECORD, dtype=torch.float64, device=samples.device)
    a = _cabi.ChainStats(samples.data_ptr(), n_iter, n_chains, size, n_sel, elem_stride, max_lag, out.data_ptr())
    _cabi.check(K.lib().omc_chain_stats(C.byref(a), K.stream_ptr()), "omc_chain_stats")
    return out


def rank_normalize(samples: torch.Tensor, elem_stride: int = 1, n_sel: int = None, pooled: bool = False) -> torch.Tensor:
    """Normal scores of the ranks (Vehtari et al. 2021) of a device store [n_iter, n_chains, size] ->
    z [n_iter, n_chains, n_sel]; ranks within each chain's series, or (pooled) over all chains of the store."""
    n_iter, n_chains, size = samples.shape
    if n_sel is None:
        n_sel = (size + elem_stride - 1) // elem_stride
    z = torch.empty(n_iter, n_chains, n_sel, dtype=torch.float64, device=samples.device)
    scratch = torch.empty(n_chains, n_sel, n_iter, dtype=torch.float64, device=samples.device)
    _cabi.check(K.lib().omc_rank_normalize(samples.data_ptr(), n_iter, n_chains, size, n_sel, elem_stride, int(bool(pooled)),
                                           scratch.data_ptr(), z.data_ptr(), K.stream_ptr()), "omc_rank_normalize")
    return z


def gather_records(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather per-chain records [n_local, n_sel, 8] over the process group -> [n_total, n_sel, 8], rank order =
    chain order.  Ranks may hold different numbers of chains (shard_chains); the counts are exchanged first."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    counts = [int(c.item()) for c in counts]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged blocks: pad every rank to the largest block, gather, drop the padding
    cmax = max(counts)
    padded = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * cmax: r * cmax + counts[r]] for r in range(world)], dim=0)


def rhat_combine(records: torch.Tensor) -> torch.Tensor:
    """records [n_chains_total, n_sel, 8] on the device -> [n_sel, 4] = split-R-hat, total ESS, grand mean, var+."""
    n_total, n_sel, _ = records.shape
    out = torch.empty(n_sel, 4, dtype=torch.float64, device=records.device)
    rec = records.contiguous()
    _cabi.check(K.lib().omc_rhat_combine(rec.data_ptr(), n_total, n_sel, out.data_ptr(), K.stream_ptr()),
                "omc_rhat_combine")
    return out


def summarize(mcmc, params=None, elem_stride=None, max_lag: int = 127, group=None, rank_normalized=None) -> dict:
    """Diagnostics of a finished (or running) `MCMC`: for every sampled parameter a dict with per-chain `records`
    [n_local, n_sel, 8], and after the all-gather `rhat`, `ess` (sum over all chains), `mean`, `var_plus` [n_sel].
    `elem_stride[param]` thins long parameters (C3: a strided subset of the field).
    rank_normalized: None = the draws themselves (Geyer ESS, plain split-R-hat); "chain" = the draws replaced by the normal
    scores of their ranks within each chain (bulk-ESS per chain, what ESS/s sums); "pooled" = ranks over all chains of a
    block (bulk-ESS and rank-normalised split-R-hat of Vehtari et al. 2021; for chains that share their target; the
    ranks are pooled per device, not across ranks).  `mean` / `var_plus` then refer to the scores."""
    if rank_normalized not in (None, "chain", "pooled"):
        raise ValueError("rank_normalized must be None, 'chain' or 'pooled'")
    out = {}
    blocks = getattr(mcmc, "_blocks", None) or [mcmc]       # a run in chain blocks keeps its plans / stores per block
    for s in mcmc.samplers:
        if params is not None and s.param not in params:
            continue
        stride = (elem_stride or {}).get(s.param, 1)
        recs = []
        for blk in blocks:
            with torch.cuda.stream(blk.stream):
                buf, eff_stride = blk.device_samples(s.param, stride)
                if rank_normalized:
                    buf, eff_stride = rank_normalize(buf, elem_stride=eff_stride, pooled=rank_normalized == "pooled"), 1
                recs.append(chain_stats(buf, elem_stride=eff_stride, max_lag=max_lag))
            blk.stream.synchronize()
        rec = recs[0] if len(recs) == 1 else torch.cat(recs, dim=0)     # chain order = block order
        allrec = gather_records(rec, group)
        comb = rhat_combine(allrec)
        out[s.param] = {"records": rec, "all_records": allrec, "rhat": comb[:, 0], "ess": comb[:, 1],
                        "mean": comb[:, 2], "var_plus": comb[:, 3], "n_chains_total": allrec.shape[0]}
    torch.cuda.synchronize()
    return out


def min_ess_per_chain(summary: dict, all_ranks: bool = False) -> torch.Tensor:
    """per chain, the minimum ESS over all scalar parameters in the summary -> [n_local] (or the chains of
    every rank, [n_total], with all_ranks=True)."""
    key = "all_records" if all_ranks else "records"
    per = [v[key][:, :, 3].min(dim=1).values for v in summary.values()]
    return torch.stack(per, dim=0).min(dim=0).values
