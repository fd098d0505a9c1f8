"""Two ranks, two GPUs, NCCL: the cross x chromosome grid of sharding.predict_sharded against a single-GPU run.
Skipped when fewer than 2 GPUs are visible (the CPU twin is tests/test_sharding_cpu.py over gloo)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(77)
    m_c = np.array([300, 170, 260, 90])
    m = int(m_c.sum())
    cM = np.concatenate([np.sort(rng.uniform(0, 120, k)) for k in m_c])
    H = (rng.random((2 * 20, m)) < 0.5).astype(np.uint8)
    add, dom = rng.standard_normal((m, 3)), rng.standard_normal((m, 3))
    sire = rng.integers(0, 20, 333).astype(np.int32)
    dam = rng.integers(0, 20, 333).astype(np.int32)
    return m_c, cM, H, add, dom, sire, dam


def _worker(rank, world, port, cross_shards, chrom_shards, q):
    import torch
    import torch.distributed as dist
    from genomicmateselectr_b200 import CrossVarEngine, sharding as S
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m_c, cM, H, add, dom, sire, dam = _problem()
    eng = CrossVarEngine(rank)
    eng.set_genmap(cM, m_c)
    eng.set_haplotypes(H)
    layout = S.GridLayout(world, cross_shards, chrom_shards)
    out, nseg = S.predict_sharded(eng, "AD", add, dom, None, sire, dam, layout, m_c, dist, device=torch.device("cuda", rank))
    if rank == 0:
        q.put((out, nseg))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("cross_shards,chrom_shards", [(2, 1), (1, 2)])
def test_two_ranks_nccl_equal_single_gpu(cross_shards, chrom_shards):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from genomicmateselectr_b200 import CrossVarEngine
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cross_shards, chrom_shards, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, nseg = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    m_c, cM, H, add, dom, sire, dam = _problem()
    with CrossVarEngine(0) as eng:
        eng.set_genmap(cM, m_c)
        eng.set_haplotypes(H)
        ref, rseg = eng.predict("AD", add, dom, None, sire, dam)
    assert np.array_equal(nseg, rseg)
    assert np.allclose(out, ref, rtol=1e-11, atol=0)
