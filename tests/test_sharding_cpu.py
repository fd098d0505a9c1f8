"""Multi-rank host logic on CPU: world_size 2 over gloo. The per-rank slabs are synthetic (this file
never touches the GPU); what is checked is the cross x chromosome grid bookkeeping and the
all-reduce / gather combination of sharding.combine()."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from genomicmateselectr_b200 import sharding as S


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 50_000):
        for parts in (1, 2, 3, 8):
            r = [S.shard_range(n, i, parts) for i in range(parts)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_balance_chromosomes():
    m_c = [1833] * 12 + [1834] * 6
    for parts in (1, 2, 3, 4, 6, 9, 18):
        r = S.balance_chromosomes(m_c, parts)
        assert r[0][0] == 0 and r[-1][1] == 18 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(b > a for a, b in r)
        w = [sum(x * x for x in m_c[a:b]) for a, b in r]
        assert max(w) <= 1.35 * sum(w) / parts or parts > 9
    assert S.balance_chromosomes([5000, 100, 100, 100], 2) == [(0, 1), (1, 4)]
    with pytest.raises(ValueError):
        S.balance_chromosomes([10, 10], 3)


def test_grid_layout():
    g = S.GridLayout(8, 4, 2)
    assert g.coords(5) == (2, 1) and g.chrom_group_ranks(2) == [4, 5] and g.gather_ranks() == [0, 2, 4, 6]
    with pytest.raises(ValueError):
        S.GridLayout(8, 3, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cross_shards, chrom_shards, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layout = S.GridLayout(world, cross_shards, chrom_shards)
    xi, ci = layout.coords(rank)
    lo, hi = S.shard_range(n_total, xi, cross_shards)
    # truth[x, p, c] = x + 0.25 p + 0.5 c; chromosome shard ci contributes the fraction (ci+1)/sum
    x = np.arange(lo, hi, dtype=np.float64)[:, None, None]
    truth = x + 0.25 * np.arange(3)[None, :, None] + 0.5 * np.arange(2)[None, None, :]
    frac = (ci + 1) / sum(range(1, chrom_shards + 1))
    out = truth * frac
    nseg = (np.arange(lo, hi) * 2 + 4).astype(np.int32) if ci == 0 else np.zeros(hi - lo, np.int32)
    o, s = S.combine(out, nseg, layout, n_total, dist)
    if rank == 0:
        q.put((o, s))
    else:
        assert o is None and s is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cross_shards,chrom_shards", [(2, 1), (1, 2)])
def test_combine_world2_gloo(cross_shards, chrom_shards):
    n_total = 11
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cross_shards, chrom_shards, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    o, s = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    x = np.arange(n_total, dtype=np.float64)[:, None, None]
    truth = x + 0.25 * np.arange(3)[None, :, None] + 0.5 * np.arange(2)[None, None, :]
    assert o.shape == (n_total, 3, 2) and np.allclose(o, truth, rtol=1e-15)
    assert np.array_equal(s, np.arange(n_total) * 2 + 4)
