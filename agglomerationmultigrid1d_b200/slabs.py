"""Slab plan of the multi-GPU variant (host-side mirror of the rule in csrc/amg1d.cu,
``alloc_level_common``): the 1-D mesh shards into contiguous element slabs, one rank per GPU.

A level is sharded while it has at least ``nranks * shard_min`` elements, divides evenly (a remainder
of ONE element - the closing vertex group of a CG level, which has n + 1 groups - goes to the last
rank) and every finer level is sharded too; the first level that is not lives on rank 0 only ("gather level"): its
right-hand side is gathered from the slabs, rank 0 runs the remaining sub-hierarchy, and the
correction is scattered back with ``ghost_depth`` ghost elements per slab edge.  ``ghost_depth`` must
be max(nPre, nPost) + 1: one halo exchange then serves a whole fused leg (S sweeps + residual); the
two-parent transfers of CG levels read ``ratio`` ghost elements more (restriction gathers from the
children of the neighbouring coarse element), so hierarchies with sharded CG levels add that.
"""


def slab_start(n_glob, nranks, r):
    return (n_glob // nranks) * r


def slab_size(n_glob, nranks, r):
    return n_glob // nranks + (n_glob % nranks if r == nranks - 1 else 0)



class LevelSlab:
    def __init__(self, n_glob, sharded, rank, nranks, ghost_depth):
        self.n_glob = n_glob
        self.sharded = sharded
        if sharded:
            self.n = slab_size(n_glob, nranks, rank)
            self.start = slab_start(n_glob, nranks, rank)
            self.gl = ghost_depth if rank > 0 else 0
            self.gr = ghost_depth if rank < nranks - 1 else 0
            self.present = True
        else:
            self.n, self.start, self.gl, self.gr = n_glob, 0, 0, 0
            self.present = rank == 0


def plan_slabs(level_sizes, ratios, rank, nranks, shard_min=8192, ghost_depth=4):
    """level_sizes[l]: elements of level l; ratios[l]: fine elements per coarse element between l, l+1.
    Returns (list of LevelSlab, gather_level); gather_level is -1 on a single rank."""
    plan, gather = [], -1
    prev_sharded = True
    for l, n in enumerate(level_sizes):
        sharded = (nranks > 1 and prev_sharded and n % nranks <= 1 and n // nranks >= shard_min)
        if sharded and l < len(ratios) and n % nranks == 0 and level_sizes[l + 1] % nranks == 0:
            nloc = n // nranks
            if nloc % ratios[l] or nloc < 2 * ghost_depth:
                raise ValueError(f"level {l}: slab of {nloc} elements is not aligned to the "
                                 f"agglomeration ratio {ratios[l]}")
        if nranks > 1 and l == 0 and not sharded:
            raise ValueError("multi-GPU: the finest level must be shardable")
        if nranks > 1 and not sharded and gather < 0:
            gather = l
        plan.append(LevelSlab(n, sharded, rank, nranks, ghost_depth))
        prev_sharded = sharded
    if nranks > 1 and gather < 0:
        raise ValueError("multi-GPU: the coarsest levels must fall below the shard threshold")
    return plan, gather
