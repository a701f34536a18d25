"""CPU emulation of the slab-sharded V-cycle schedule of csrc/amg1d.cu (enqueue_vcycle, leg_down,
leg_up, op_halo, op_gather_rhs, op_scatter_sol) with numpy blocks and torch.distributed (gloo).

Each fused leg is emulated the way the kernels compute it: S Jacobi sweeps over the slab extended by
its ghost elements with zeros beyond (so values near the ghost edge go stale exactly as in the
kernel's window), then only the owned elements are emitted.  If the schedule exchanged too little or
too late, the owned results would differ from the single-rank run."""
import numpy as np
import torch
import torch.distributed as dist

from agglomerationmultigrid1d_b200.slabs import plan_slabs, slab_size, slab_start


class Comm:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def sendrecv(self, peer, send):
        recv = torch.empty_like(send)
        if self.rank < peer:
            dist.send(send, peer); dist.recv(recv, peer)
        else:
            dist.recv(recv, peer); dist.send(send, peer)
        return recv


def halo(comm, v, n, gd):
    """v: (gl + n + gr, m) array of a slab with its ghosts; fills the ghost rows from the neighbours."""
    gl = gd if comm.rank > 0 else 0
    if comm.rank > 0:
        got = comm.sendrecv(comm.rank - 1, torch.from_numpy(v[gl:gl + gd].copy()))
        v[0:gd] = got.numpy()
    if comm.rank < comm.world - 1:
        got = comm.sendrecv(comm.rank + 1, torch.from_numpy(v[gl + n - gd:gl + n].copy()))
        v[gl + n:gl + n + gd] = got.numpy()


def sweep(ops, b, x, alpha, zero):
    lo, di, up, dinv = ops
    if zero:
        r = b.copy()
    else:
        xl = np.vstack([np.zeros((1, x.shape[1])), x[:-1]])
        xr = np.vstack([x[1:], np.zeros((1, x.shape[1]))])
        r = b - (np.einsum("eij,ej->ei", lo, xl) + np.einsum("eij,ej->ei", di, x)
                 + np.einsum("eij,ej->ei", up, xr))
    return x + alpha * np.einsum("eij,ej->ei", dinv, r)


def residual(ops, b, x):
    lo, di, up, _ = ops
    xl = np.vstack([np.zeros((1, x.shape[1])), x[:-1]])
    xr = np.vstack([x[1:], np.zeros((1, x.shape[1]))])
    return b - (np.einsum("eij,ej->ei", lo, xl) + np.einsum("eij,ej->ei", di, x)
                + np.einsum("eij,ej->ei", up, xr))


def transfer_of(U, l):
    """(parent, P0, P1) of transfer l with explicit global arrays: fine element e receives
    P0[e] x_c[parent[e]] + P1[e] x_c[parent[e] + 1]  (P1 None for single-parent transfers)."""
    if hasattr(U, "cg_orders"):
        return U.transfer_blocks(l)
    P, ratio = U.transfers[l]
    e = np.arange(U.levels[l].n)
    return e // ratio, P[e % ratio], None


def level_ratios(U):
    if hasattr(U, "cg_orders"):
        return [t["ratio"] for t in U.cg_transfers] + [r for (_, r) in U.transfers]
    return [r for (_, r) in U.transfers]


def smoother_inverse(U, l, di):
    """Block Jacobi (DG-type levels) or point Jacobi (CG levels) as explicit inverse blocks."""
    if getattr(U.levels[l], "is_cg", False):
        d = np.zeros_like(di)
        idx = np.arange(di.shape[1])
        d[:, idx, idx] = 1.0 / di[:, idx, idx]
        return d
    return np.linalg.inv(di)


def vcycle(U, x0_glob, b_glob, rank, world, nPre=3, nPost=3, alpha=2.0 / 3.0, shard_min=16, gd=4):
    """Returns this rank's owned slab of x after one V-cycle (the whole vector when world == 1).
    Works for DG-first hierarchies (single-parent transfers, slabs of elements) and CG-first ones
    (two-parent transfers across the slab edges, slabs of vertex groups with the closing group on the
    last rank); gd must be max(nPre, nPost) + 1 (+ the ratio of the two-parent transfers)."""
    comm = Comm(rank, world)
    nL = len(U.levels)
    sizes = [lv.n for lv in U.levels]
    ratios = level_ratios(U)
    plan, g = plan_slabs(sizes, ratios, rank, world, shard_min=shard_min, ghost_depth=gd)
    ops, b, x = [None] * nL, [None] * nL, [None] * nL
    for l, sl in enumerate(plan):
        if not sl.present and l != g:
            continue
        m = U.levels[l].m
        lo, di, up = U.level_blocks(l)
        a, e = sl.start - sl.gl, sl.start + sl.n + sl.gr
        if sl.present:
            ops[l] = (lo[a:e], di[a:e], up[a:e], smoother_inverse(U, l, di[a:e]))
        b[l] = np.zeros((e - a, m))
        x[l] = np.zeros((e - a, m))
    own = lambda l: slice(plan[l].gl, plan[l].gl + plan[l].n)             # noqa: E731
    m0 = U.levels[0].m
    s0 = plan[0]
    b[0][own(0)] = b_glob.reshape(-1, m0)[s0.start:s0.start + s0.n]
    x[0][own(0)] = x0_glob.reshape(-1, m0)[s0.start:s0.start + s0.n]
    if s0.sharded:
        halo(comm, b[0], s0.n, gd)
    # proxy slab of the gather level on ranks > 0
    if world > 1 and rank > 0:
        ng = slab_size(sizes[g], world, rank)
        mg = U.levels[g].m
        b[g] = np.zeros((ng, mg))
        x[g] = np.zeros((gd + ng + (gd if rank < world - 1 else 0), mg))
    # ---- down ----
    for l in range(nL - 1):
        sl = plan[l]
        if not sl.present:
            break
        parent, P0, P1 = transfer_of(U, l)
        zero = l > 0
        if sl.sharded and not zero:
            halo(comm, x[l], sl.n, gd)
        xx = np.zeros_like(x[l]) if zero else x[l].copy()
        for s in range(nPre):
            xx = sweep(ops[l], b[l], xx, alpha, zero and s == 0)
        r = residual(ops[l], b[l], xx)
        x[l][own(l)] = xx[own(l)]
        # restriction into this rank's coarse elements, from the slab extended by its ghosts (stale
        # values near the ghost edge included, exactly what the kernel's window would read)
        nxt = plan[l + 1]
        c_lo = slab_start(sizes[l + 1], world, rank) if sl.sharded else 0
        c_n = slab_size(sizes[l + 1], world, rank) if sl.sharded else sizes[l + 1]
        eg = (sl.start - sl.gl) + np.arange(r.shape[0])
        rc = np.zeros((c_n, U.levels[l + 1].m))
        for P, off in ((P0, 0), (P1, 1)):
            if P is None:
                continue
            t = np.einsum("eij,ei->ej", P[eg], r)
            k = parent[eg] + off - c_lo
            ok = (k >= 0) & (k < c_n)
            np.add.at(rc, k[ok], t[ok])
        if sl.sharded:
            halo(comm, x[l], sl.n, gd)
            if nxt.sharded:
                b[l + 1][own(l + 1)] = rc
                halo(comm, b[l + 1], nxt.n, gd)
            else:                                           # gather to rank 0
                parts = [None] * world
                dist.gather_object(rc, parts if rank == 0 else None, dst=0)
                if rank == 0:
                    b[l + 1][:] = np.vstack(parts)
        else:
            b[l + 1][own(l + 1)] = rc
    # ---- coarsest ----
    if plan[nL - 1].present:
        lo, di, up = U.level_blocks(nL - 1)
        n, m = di.shape[0], di.shape[1]
        A = np.zeros((n * m, n * m))
        for e in range(n):
            A[e * m:(e + 1) * m, e * m:(e + 1) * m] = di[e]
            if e > 0:
                A[e * m:(e + 1) * m, (e - 1) * m:e * m] = lo[e]
            if e < n - 1:
                A[e * m:(e + 1) * m, (e + 1) * m:(e + 2) * m] = up[e]
        x[nL - 1] = np.linalg.solve(A, b[nL - 1].ravel()).reshape(n, m)
    # ---- up ----
    for l in range(nL - 2, -1, -1):
        sl, nxt = plan[l], plan[l + 1]
        if not sl.present:
            continue
        parent, P0, P1 = transfer_of(U, l)
        if sl.sharded and not nxt.sharded:                  # scatter rank 0 -> slabs (+ ghosts)
            if rank == 0:
                objs = []
                for r_ in range(world):
                    c0, cn = slab_start(sizes[l + 1], world, r_), slab_size(sizes[l + 1], world, r_)
                    objs.append(x[l + 1][max(0, c0 - gd):c0 + cn + (gd if r_ < world - 1 else 0)])
            else:
                objs = None
            got = [None]
            dist.scatter_object_list(got, objs, src=0)
            xc, c_first = (x[l + 1], 0) if rank == 0 else (got[0], slab_start(sizes[l + 1], world, rank) - gd)
        else:
            xc, c_first = x[l + 1], nxt.start - nxt.gl
        a = sl.start - sl.gl
        eg = a + np.arange(x[l].shape[0])
        xx = x[l].copy()
        for P, off in ((P0, 0), (P1, 1)):
            if P is None:
                continue
            par = parent[eg] + off - c_first
            ok = (par >= 0) & (par < xc.shape[0])
            xx[ok] = xx[ok] + np.einsum("eij,ej->ei", P[eg[ok]], xc[par[ok]])
        for s in range(nPost):
            xx = sweep(ops[l], b[l], xx, alpha, False)
        x[l][own(l)] = xx[own(l)]
        if sl.sharded and l > 0:
            halo(comm, x[l], sl.n, gd)
    return x[0][own(0)].ravel()
