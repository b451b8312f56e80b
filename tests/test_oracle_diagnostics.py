"""CPU: the numpy diagnostics oracle against closed forms (parity unpinned — the reference has no ESS / R-hat), the
chain sharding arithmetic, and the all-gather of per-chain records over a world_size-2 gloo group."""

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ess_of_ar1_matches_theory():
    from oracle import diagnostics as D

    rng = np.random.default_rng(0)
    N, C = 4000, 6
    for rho in (0.0, 0.5, 0.9):
        x = np.zeros((N, C, 1))
        e = rng.standard_normal((N, C))
        x[0, :, 0] = e[0]
        for t in range(1, N):
            x[t, :, 0] = rho * x[t - 1, :, 0] + np.sqrt(1 - rho ** 2) * e[t]
        st = D.chain_stats(x)
        theory = N * (1 - rho) / (1 + rho)
        assert abs(st[:, 0, 3].mean() / theory - 1) < 0.2, (rho, st[:, 0, 3].mean(), theory)
        np.testing.assert_allclose(st[:, 0, 1], x[:, :, 0].mean(axis=0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(st[:, 0, 2], x[:, :, 0].var(axis=0, ddof=1), rtol=1e-12)
        r = D.rhat_combine(st)
        assert abs(r[0, 0] - 1) < 0.05 and abs(r[0, 1] - st[:, 0, 3].sum()) < 1e-9


def test_rhat_flags_chains_that_disagree():
    from oracle import diagnostics as D

    rng = np.random.default_rng(1)
    x = rng.standard_normal((500, 4, 2))
    x[:, 0, 1] += 3.0   # one chain of parameter 1 sits elsewhere
    r = D.rhat_combine(D.chain_stats(x))
    assert r[0, 0] < 1.05 and r[1, 0] > 1.5


def test_rank_normalized_diagnostics_handle_heavy_tails():
    """Bulk-ESS / rank-normalised split-R-hat (Vehtari et al. 2021): i.i.d. Cauchy chains have no variance, the plain
    estimators are erratic, the rank-normalised ones see S independent draws and R-hat = 1; a shifted chain is flagged."""
    from oracle import diagnostics as D

    rng = np.random.default_rng(2)
    N, C = 1000, 4
    x = rng.standard_cauchy((N, C, 1))
    z = D.rank_normalize(x, pooled=True)
    assert abs(z.mean()) < 1e-12 and abs(z.std() - 1) < 0.01            # normal scores of a permutation of 1..S
    st = D.chain_stats(z)
    r = D.rhat_combine(st)
    assert abs(r[0, 0] - 1) < 0.01 and r[0, 1] > 0.7 * N * C
    x[:, 0, 0] += 5.0
    r2 = D.rhat_combine(D.chain_stats(D.rank_normalize(x, pooled=True)))
    assert r2[0, 0] > 1.2
    zc = D.rank_normalize(x, pooled=False)                               # per chain: every series is the same score set
    np.testing.assert_allclose(np.sort(zc[:, 0, 0]), np.sort(zc[:, 1, 0]), rtol=0, atol=1e-14)


def test_shard_chains_partitions_every_chain_once():
    from openmcmc_b200.diagnostics import shard_chains

    for total in (1, 7, 64, 65536, 8191):
        for world in (1, 2, 3, 8):
            blocks = [shard_chains(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            for (o0, n0), (o1, _) in zip(blocks, blocks[1:]):
                assert o1 == o0 + n0
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["OMC_ROOT"])
from openmcmc_b200.diagnostics import shard_chains, gather_records
from oracle import diagnostics as D
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = int(os.environ["OMC_TOTAL"])
rng = np.random.default_rng(5)
x = rng.standard_normal((300, total, 3)).cumsum(axis=0) * 0.05 + rng.standard_normal((300, total, 3))
off, n_local = shard_chains(total, rank, world)
local = torch.from_numpy(D.chain_stats(x[:, off:off + n_local]))     # each rank summarises only its own chains
allrec = gather_records(local).numpy()
ref = D.chain_stats(x)
np.testing.assert_allclose(allrec, ref, rtol=1e-12, atol=1e-12)
comb = D.rhat_combine(allrec)
np.testing.assert_allclose(comb, D.rhat_combine(ref), rtol=1e-12)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def _run_world2(total):
    env = dict(os.environ, OMC_ROOT=ROOT, OMC_TOTAL=str(total), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + os.getpid() % 400), "-c", WORKER]
    # torch.distributed.run wants a script path: write the worker next to the test's temp dir
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "worker.py")
        open(path, "w").write(WORKER)
        cmd[-2:] = [path]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == 2


def test_all_gather_of_chain_records_world_size_2_gloo_even_split():
    _run_world2(8)


def test_all_gather_of_chain_records_world_size_2_gloo_ragged_split():
    _run_world2(7)
