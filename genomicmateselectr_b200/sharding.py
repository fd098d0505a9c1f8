"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

The path shards on two independent axes:
  * by cross       - crosses are independent; every rank holds all LD-decay tiles + haplotypes and
                     takes a contiguous slice of the cross list. No data-path collective; only the
                     final gather of the per-rank result slabs.
  * by chromosome  - every variance and Nsegsnps is a sum over chromosomes; a rank that only owns some
                     chromosomes produces PARTIAL slabs which are added with one all-reduce(sum, FP64)
                     inside its chromosome group (fixed ring order => reproducible for a fixed layout).
A `world = cross_shards x chrom_shards` grid combines both. Rank r has chromosome index r % chrom_shards
and cross index r // chrom_shards."""
from dataclasses import dataclass

import numpy as np


def shard_range(n, index, parts):
    """Contiguous, balanced slice [lo, hi) of n items for shard `index` of `parts` (first n % parts get +1)."""
    base, extra = divmod(int(n), int(parts))
    lo = index * base + min(index, extra)
    return lo, lo + base + (1 if index < extra else 0)


def balance_chromosomes(m_c, parts):
    """Contiguous chromosome ranges [(c_begin, c_end)] whose sum of m_c^2 (the contraction work) is as
    even as a contiguous split allows. Every range is non-empty (parts <= number of chromosomes)."""
    m_c = np.asarray(m_c, dtype=np.float64)
    C = len(m_c)
    if not 1 <= parts <= C:
        raise ValueError("chromosome shards must be between 1 and the number of chromosomes")
    w = m_c ** 2
    cum = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for p in range(1, parts):
        target = cum[-1] * p / parts
        lo, hi = bounds[-1] + 1, C - (parts - p)
        c = int(np.clip(np.searchsorted(cum, target), lo, hi))
        if c - 1 >= lo and abs(cum[c - 1] - target) <= abs(cum[c] - target):
            c -= 1
        bounds.append(c)
    bounds.append(C)
    return [(bounds[i], bounds[i + 1]) for i in range(parts)]


@dataclass
class GridLayout:
    world: int
    cross_shards: int
    chrom_shards: int

    def __post_init__(self):
        if self.cross_shards * self.chrom_shards != self.world:
            raise ValueError("cross_shards x chrom_shards must equal the world size")

    def coords(self, rank):
        return rank // self.chrom_shards, rank % self.chrom_shards      # (cross index, chromosome index)

    def chrom_group_ranks(self, cross_index):
        return [cross_index * self.chrom_shards + c for c in range(self.chrom_shards)]

    def gather_ranks(self):
        return [x * self.chrom_shards for x in range(self.cross_shards)]   # chromosome index 0 of every cross shard


_GROUPS = {}


def grid_groups(layout, dist):
    """Process groups of a grid layout, created ONCE per (world, layout) and reused by every later call (new_group is a
    collective over the whole world and must not sit in a loop). Returns (chromosome groups per cross shard, gather
    group of the chromosome-index-0 ranks, cross-axis groups per chromosome index)."""
    key = (dist.get_world_size(), layout.cross_shards, layout.chrom_shards, dist.get_backend())
    g = _GROUPS.get(key)
    if g is None:
        # every rank creates every group, in the same order
        chrom_groups = [dist.new_group(layout.chrom_group_ranks(x)) for x in range(layout.cross_shards)] \
            if layout.chrom_shards > 1 else None
        gather_group = dist.new_group(layout.gather_ranks()) if layout.chrom_shards > 1 else None
        cross_groups = [dist.new_group([x * layout.chrom_shards + c for x in range(layout.cross_shards)])
                        for c in range(layout.chrom_shards)] if (layout.chrom_shards > 1 and layout.cross_shards > 1) else None
        g = _GROUPS[key] = (chrom_groups, gather_group, cross_groups)
    return g


def reset_groups():
    """Forget cached groups (call after dist.destroy_process_group())."""
    _GROUPS.clear()


def combine(out, nseg, layout, n_total, dist, device=None):
    """Combine per-rank slabs: all-reduce over the chromosome axis, gather over the cross axis.
    `out` [n_local, npairs, ncomp] float64 and `nseg` [n_local] int32 are NumPy arrays; `dist` is
    torch.distributed (already initialised; gloo on CPU or nccl on `device`). Returns (out, nseg) for
    all `n_total` crosses on rank 0 and (None, None) elsewhere."""
    import torch
    rank = dist.get_rank()
    xi, ci = layout.coords(rank)
    t_out = torch.from_numpy(np.ascontiguousarray(out))
    t_seg = torch.from_numpy(np.ascontiguousarray(nseg))
    if device is not None:
        t_out, t_seg = t_out.to(device), t_seg.to(device)
    chrom_groups, gather_group, _ = grid_groups(layout, dist)
    if layout.chrom_shards > 1:
        dist.all_reduce(t_out, op=dist.ReduceOp.SUM, group=chrom_groups[xi])
        dist.all_reduce(t_seg, op=dist.ReduceOp.SUM, group=chrom_groups[xi])
    res = (None, None)
    if ci == 0:
        sizes = [shard_range(n_total, x, layout.cross_shards) for x in range(layout.cross_shards)]
        nmax = max(hi - lo for lo, hi in sizes)
        pad_out = torch.zeros((nmax,) + tuple(t_out.shape[1:]), dtype=t_out.dtype, device=t_out.device)
        pad_seg = torch.zeros((nmax,), dtype=t_seg.dtype, device=t_seg.device)
        pad_out[: t_out.shape[0]] = t_out
        pad_seg[: t_seg.shape[0]] = t_seg
        if layout.cross_shards > 1:
            bufs_o = [torch.empty_like(pad_out) for _ in sizes] if rank == 0 else None
            bufs_s = [torch.empty_like(pad_seg) for _ in sizes] if rank == 0 else None
            dist.gather(pad_out, bufs_o, dst=0, group=gather_group)
            dist.gather(pad_seg, bufs_s, dst=0, group=gather_group)
        else:
            bufs_o, bufs_s = [pad_out], [pad_seg]
        if rank == 0:
            o = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs_o, sizes)]).cpu().numpy()
            s = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs_s, sizes)]).cpu().numpy()
            res = (o, s)
    return res


def stage_local(engine, model, add, dom, asub, sire, dam, layout, m_c, dist):
    """Host -> device for this rank's part of a cross x chromosome grid: chromosome shard, cross range [lo, hi) of the
    FULL cross list, and the split of the per-parent stage over the ranks that hold the same chromosomes."""
    rank = dist.get_rank() if layout.world > 1 else 0
    xi, ci = layout.coords(rank)
    if layout.chrom_shards > 1:
        c0, c1 = balance_chromosomes(m_c, layout.chrom_shards)[ci]
        engine.set_chromosome_shard(c0, c1)
    lo, hi = shard_range(len(sire), xi, layout.cross_shards)
    engine.set_parent_shard(xi if layout.cross_shards > 1 else 0, layout.cross_shards)
    engine.stage_range(model, add, dom, asub, sire, dam, lo, hi)
    return lo, hi


def compute_local(engine, layout, dist, sync=True, reduce_chromosomes=False):
    """Device part on the staged inputs: per-parent stage (this rank's slice), ONE all-gather of the table slices over
    the ranks with the same chromosome index, per-cross stage; optionally the all-reduce(sum) of the partial result
    slabs over the chromosome axis, in place in the engine's device buffers."""
    engine.compute_parents()
    if layout.cross_shards > 1:
        _, _, cross_groups = grid_groups(layout, dist)
        ci = layout.coords(dist.get_rank())[1]
        engine.exchange_parents(dist, cross_groups[ci] if cross_groups is not None else None)
    engine.compute_crosses(sync=sync and not (reduce_chromosomes and layout.chrom_shards > 1))
    if reduce_chromosomes and layout.chrom_shards > 1:
        import torch
        chrom_groups, _, _ = grid_groups(layout, dist)
        xi = layout.coords(dist.get_rank())[0]
        out, nseg = engine.device_outputs()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=chrom_groups[xi])
        dist.all_reduce(nseg, op=dist.ReduceOp.SUM, group=chrom_groups[xi])
        if sync:
            torch.cuda.synchronize()


def run_local(engine, model, add, dom, asub, sire, dam, layout, m_c, dist, sync=True):
    """stage_local + compute_local. Every rank calls this with the FULL cross list. Returns (lo, hi)."""
    lo, hi = stage_local(engine, model, add, dom, asub, sire, dam, layout, m_c, dist)
    compute_local(engine, layout, dist, sync=sync)
    return lo, hi


def predict_sharded(engine, model, add, dom, asub, sire, dam, layout, m_c, dist, device=None):
    """The hot path on a cross x chromosome grid of ranks. Every rank calls this with the FULL cross list."""
    run_local(engine, model, add, dom, asub, sire, dam, layout, m_c, dist)
    out, nseg = engine.fetch()
    return combine(out, nseg, layout, len(sire), dist, device)
