"""Scalable hierarchy set-up for uniform meshes (2^20 .. 2^26+ elements).

The general path (mesh_hierarchy.py) follows the reference literally: global sparse matrices and
sparse triple products.  At 2^26 elements that is tens of GB of host memory, so for *uniform* meshes
this module builds the same DG -> ... -> DG -> agglomerated hierarchy in a translation-invariant
"pattern" form: on every level only the first / last ``NB`` element block sets (which feel the
boundary conditions) and one interior block set are distinct.  Each coarsening step

    G, D, C  <-  L' (G, D, C) L ,     A = C - D (M \\ G)              (src/mesh_heirarchy.jl:75-106)

is carried out with the same scipy algebra as the general path, on a small *virtual window* of the
level (first NV/2 and last NV/2 elements stitched together), and compressed back to a pattern; the
window is exact because every operator is block tridiagonal and every transfer is element local.
Once a level has at most ``EXPLICIT_BELOW`` elements the hierarchy continues with explicit arrays.

Transfer blocks and agglomerated mass matrices use the closed forms that the reference's quadrature
sums reduce to on a uniform mesh (they integrate polynomials of degree <= 2 exactly):
    aggdg_dg      (src/interpolation.jl:270-292)  rows [1, ((2a+1-r) + xi_i)/r]  for child a of r
    aggdg_aggdg   (src/interpolation.jl:226-264)  [[1, (2a+1-r)/r], [0, 1/r]]
    mass          (src/agglomerated_dg_mesh.jl:441-453)  H diag(1, 1/3)
tests/test_uniform_patterns.py checks the result against the general path block by block.
"""
import math

import numpy as np
import scipy.sparse as sp

from . import blocks as blk
from .device import DeviceHierarchy
from .dg_mesh import DgMesh, dg_flux_operators, eval_func
from .meshes import create_uniform_mesh, set_boundary
from .reference_element import ReferenceElement, evaluate_nodal_basis_fun

NB = 4                 # explicit head / tail elements per level
NV = 32                # virtual window size (elements) used for one coarsening step
EXPLICIT_BELOW = 32    # levels with at most this many elements are kept explicit


class Pattern:
    """Block-tridiagonal operator with head / interior / tail structure.
    lo, di, up: (NB + 1 + NB, mr, mc) block arrays in (e, i, j) order."""

    def __init__(self, lo, di, up):
        self.lo, self.di, self.up = lo, di, up

    def expand(self, n):
        """Explicit (n, mr, mc) arrays for a level with n >= 2 NB + 1 elements."""
        idx = pattern_index(n)
        return self.lo[idx], self.di[idx], self.up[idx]


def pattern_index(n):
    if n < 2 * NB + 1:
        raise ValueError("level too small for a pattern")
    idx = np.full(n, NB, dtype=np.int64)
    idx[:NB] = np.arange(NB)
    idx[n - NB:] = NB + 1 + np.arange(NB)
    return idx


def compress(lo, di, up, what=""):
    """Explicit window arrays -> Pattern, checking that the interior is translation invariant."""
    n = di.shape[0]
    mid = n // 2
    for name, a in (("lo", lo), ("di", di), ("up", up)):
        inner = a[NB:n - NB]
        scale = max(np.abs(a).max(), 1e-300)
        dev = np.abs(inner - a[mid]).max() / scale
        if dev > 1e-11:
            raise RuntimeError(f"{what}.{name}: interior blocks are not translation invariant "
                               f"(relative deviation {dev:.2e})")
    pick = np.concatenate([np.arange(NB), [mid], np.arange(n - NB, n)])
    return Pattern(lo[pick].copy(), di[pick].copy(), up[pick].copy())


def _csc(lo, di, up):
    n, mr, mc = di.shape
    rs = np.arange(n * mr).reshape(n, mr)
    cs = np.arange(n * mc).reshape(n, mc)
    rows, cols, vals = [], [], []
    for a, off in ((lo, -1), (di, 0), (up, 1)):
        e = np.arange(n)
        ok = (e + off >= 0) & (e + off < n)
        rows.append(np.repeat(rs[e[ok]][:, :, None], mc, axis=2).ravel())
        cols.append(np.repeat(cs[e[ok] + off][:, None, :], mr, axis=1).ravel())
        vals.append(a[ok].ravel())
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(n * mr, n * mc))


def _blocks(A, n, m):
    return blk.csc_to_blocks(A, np.arange(n * m, dtype=np.int64).reshape(n, m))


def _transfer_csc(P, n_fine, ratio):
    """Explicit prolongation for n_fine elements from its periodic blocks P (ratio, mf, mc)."""
    _, mf, mc = P.shape
    e = np.arange(n_fine)
    rows = np.repeat((e[:, None] * mf + np.arange(mf))[:, :, None], mc, axis=2).ravel()
    cols = np.repeat(((e // ratio)[:, None] * mc + np.arange(mc))[:, None, :], mf, axis=1).ravel()
    vals = P[e % ratio].ravel()
    return sp.csc_matrix((vals, (rows, cols)), shape=(n_fine * mf, (n_fine // ratio) * mc))


def dg_dg_blocks(p_low, p_high):
    low, high = ReferenceElement(p_low), ReferenceElement(p_high)
    return evaluate_nodal_basis_fun(low.mBasisFunCoeff, high.mNodesX)[None]        # (1, mf, mc)


def aggdg_dg_blocks(pAgg, p_base, ratio):
    xi = ReferenceElement(p_base).mNodesX
    P = np.zeros((ratio, len(xi), pAgg + 1))
    P[:, :, 0] = 1.0
    if pAgg == 1:
        for a in range(ratio):
            P[a, :, 1] = ((2 * a + 1 - ratio) + xi) / ratio
    return P


def aggdg_aggdg_blocks(pAgg, ratio):
    P = np.zeros((ratio, pAgg + 1, pAgg + 1))
    P[:, 0, 0] = 1.0
    if pAgg == 1:
        for a in range(ratio):
            P[a, 0, 1] = (2 * a + 1 - ratio) / ratio
            P[a, 1, 1] = 1.0 / ratio
    return P


def agg_mass_block(pAgg, H):
    return np.array([[H]]) if pAgg == 0 else np.diag([H, H / 3.0])


class UniformLevel:
    """One level in pattern or explicit form: operators A (and G, D, C for the next projection)."""

    def __init__(self, n, m):
        self.n, self.m = n, m
        self.explicit = n <= EXPLICIT_BELOW
        self.ops = {}          # name -> Pattern or (lo, di, up) explicit arrays
        self.mass = None       # single (m, m) block (uniform mesh: every element has the same one)

    def window(self, name):
        """Explicit arrays of operator `name` on the virtual window (or on the level itself)."""
        o = self.ops[name]
        if self.explicit:
            return o
        return o.expand(self.window_size())

    def window_size(self):
        return self.n if self.explicit else NV


def _level_from_window(level, G, D, C, Mblock):
    """Set A = C - D (M \\ G) and store G, D, C, A on `level` from explicit window matrices."""
    nw, m = level.window_size(), level.m
    Minv = sp.kron(sp.identity(nw), np.linalg.inv(Mblock)).tocsc()
    A = (C - D @ (Minv @ G)).tocsc()
    for name, X in (("G", G), ("D", D), ("C", C), ("A", A)):
        lo, di, up = _blocks(X, nw, m)
        level.ops[name] = (lo, di, up) if level.explicit else compress(lo, di, up, name)
    level.mass = Mblock


def extend_chain(levels, transfers, steps, n_cur, H, p_cur, is_agg, pAgg):
    """Append the DG-type levels below levels[-1] (which must carry G, D, C): p-coarsened DG levels
    ("dg", p) and agglomerated levels ("agg", factor), each by L'(G,D,C)L and A = C - D (M \\ G) on a
    virtual window (src/mesh_heirarchy.jl:75-106).  transfers gets (P blocks, ratio) per step."""
    for kind, val in steps:
        fine = levels[-1]
        if kind == "dg":
            P, ratio, m_c = dg_dg_blocks(val, p_cur), 1, val + 1
            Mblock = (H / 2.0) * ReferenceElement(val).mMassMatrix
            p_cur = val
        else:
            ratio, m_c = val, pAgg + 1
            if n_cur % ratio:
                raise ValueError("agglomeration factor does not divide the element count")
            P = aggdg_aggdg_blocks(pAgg, ratio) if is_agg else aggdg_dg_blocks(pAgg, p_cur, ratio)
            H *= ratio
            Mblock = agg_mass_block(pAgg, H)
            is_agg = True
        n_next = n_cur // ratio
        coarse = UniformLevel(n_next, m_c)
        # window of the fine level on which this coarsening step is evaluated
        if coarse.explicit and not fine.explicit:
            nwf = n_cur                              # last pattern level: expand it fully
            ops = {k: fine.ops[k].expand(nwf) for k in ("G", "D", "C")}
        elif fine.explicit:
            nwf = n_cur
            ops = {k: fine.ops[k] for k in ("G", "D", "C")}
        else:
            nwf = NV * ratio
            ops = {k: fine.ops[k].expand(nwf) for k in ("G", "D", "C")}
        L = _transfer_csc(P, nwf, ratio)
        proj = {k: (L.T @ _csc(*ops[k]) @ L).tocsc() for k in ("G", "D", "C")}
        assert nwf // ratio == coarse.window_size()
        _level_from_window(coarse, proj["G"], proj["D"], proj["C"], Mblock)
        levels.append(coarse)
        transfers.append((P, ratio))
        n_cur = n_next
    return n_cur, H


class UniformDgHierarchy:
    """DG(p_0) -> DG(p_1) -> ... -> agglomerated(pAgg) levels on a uniform mesh of n elements.

    dg_orders: e.g. [3, 1]; agg_factors: e.g. [2] * 26.  Boundary kinds as in the reference scripts
    (Neumann left, Dirichlet right by default).  The operators are those of
    ``MeshHierarchy(meshes, bdConds, A, G, D, C; nDG, nAgg)`` built by the general path.
    """

    def __init__(self, n, dg_orders, agg_factors, pAgg=1, xin=0.0, xout=1.0, CDir=None,
                 bc_kinds=("neu", "dir")):
        self.n = int(n)
        self.dg_orders = list(dg_orders)
        self.agg_factors = list(agg_factors)
        self.pAgg = pAgg
        self.xin, self.xout = float(xin), float(xout)
        self.h = (self.xout - self.xin) / self.n
        self.CDir = 1000.0 * n if CDir is None else float(CDir)
        self.bc_kinds = tuple(bc_kinds)
        if not self.dg_orders:
            raise ValueError("At least one DG mesh required.")
        self.levels = []
        self.transfers = []          # (P blocks (ratio, mf, mc), ratio)
        self._build()

    # ---- level 0 from a literal small replica ----------------------------------------------------
    def _replica(self, nw):
        mesh = create_uniform_mesh(nw, self.xin, self.xin + nw * self.h)
        bd = set_boundary(mesh, self.xin, self.xin + nw * self.h,
                          [(self.bc_kinds[0], 0.0), (self.bc_kinds[1], 0.0)])
        return mesh, bd

    def _build(self):
        p0 = self.dg_orders[0]
        lv = UniformLevel(self.n, p0 + 1)
        nw = lv.window_size()
        mesh, bd = self._replica(nw)
        dgm = DgMesh(mesh, p0)
        G, D, C = dg_flux_operators(dgm, mesh, bd, self.CDir)
        _level_from_window(lv, G, D, C, dgm.mMassMatrix.mBlocks[nw // 2])
        self.levels.append(lv)
        steps = [("dg", p) for p in self.dg_orders[1:]] + [("agg", f) for f in self.agg_factors]
        extend_chain(self.levels, self.transfers, steps, self.n, self.h, p0, False, self.pAgg)

    # ---- explicit blocks of any level (for tests / small levels) -----------------------------------
    def level_blocks(self, l, name="A"):
        lv = self.levels[l]
        return lv.ops[name] if lv.explicit else lv.ops[name].expand(lv.n)

    # ---- right-hand side b = f - D (M \ r) on level 0 (src/dg_mesh.jl:342-457) ---------------------
    def rhs(self, func, bc_values, chunk=1 << 20, elem_range=None):
        """b on level 0; ``elem_range = (e_begin, e_end)`` returns only that slab (multi-GPU)."""
        p0 = self.dg_orders[0]
        m = p0 + 1
        ref = ReferenceElement(p0)
        n, h = self.n, self.h
        lo_e, hi_e = (0, n) if elem_range is None else elem_range
        b = np.empty((hi_e - lo_e) * m)
        W = ref.mGaussQuadWeights[:, None] * ref.mBasisGQFunVal          # (nq, m)
        for e0 in range(lo_e, hi_e, chunk):
            e1 = min(hi_e, e0 + chunk)
            i = np.arange(e0, e1, dtype=np.float64)
            xl = self.xin + (i / n) * (self.xout - self.xin)
            xr = self.xin + ((i + 1) / n) * (self.xout - self.xin)
            hh, xc = xr - xl, (xl + xr) / 2.0
            xq = xc[:, None] + (hh / 2.0)[:, None] * ref.mGaussQuadNodes[None, :]
            b[(e0 - lo_e) * m:(e1 - lo_e) * m] = ((hh / 2.0)[:, None] * (eval_func(func, xq) @ W)).ravel()
        for slot, val, _ in self.rhs_boundary_fixes(bc_values):      # (all "add")
            if lo_e * m <= slot < hi_e * m:
                b[slot - lo_e * m] += val
        return b

    def rhs_boundary_fixes(self, bc_values):
        """The boundary terms of b = f - D (M \\ r) (src/dg_mesh.jl:366-457: penalty and flux terms at the two end
        elements) as a list of (global level-0 slot, value, "add") - a handful of entries, shared by the host
        path above and the device-side assembly."""
        p0 = self.dg_orders[0]
        m = p0 + 1
        n = self.n
        lv = self.levels[0]
        Dlo, Ddi, Dup = lv.ops["D"] if lv.explicit else lv.ops["D"].expand(min(n, NV))
        Minv = np.linalg.inv(lv.mass)
        e1, e2 = 0, (1 if p0 >= 1 else 0)
        fixes = []

        def add(el, vec):
            fixes.extend((el * m + i, float(v), "add") for i, v in enumerate(vec) if v != 0.0)

        for side, el, loc, sgn in ((0, 0, e1, -1.0), (1, n - 1, e2, 1.0)):
            val = bc_values[side]
            unit = np.zeros(m)
            unit[loc] = 1.0
            if self.bc_kinds[side] == "dir":
                add(el, self.CDir * val * unit)
                s = Minv @ (sgn * val * unit)
                w = el if side == 0 else Ddi.shape[0] - 1            # same block in the window
                add(el, -(Ddi[w] @ s))
                if side == 1 and n > 1:
                    add(el - 1, -(Dup[w - 1] @ s))
            else:
                add(el, sgn * val * unit)
        return fixes

    def device_rhs(self, terms, bc_values, dev=None):
        """The same right-hand side assembled ON THE DEVICE (amg1d_dev_assemble_rhs; every rank its own slab):
        func(x) = sum of terms (kind, coef, pow, w, phi), e.g. [("cos", w * w, 0, w, 0.0)].  Leaves b on level 0 and
        x = 0; nothing but the quadrature tables and the boundary entries crosses PCIe."""
        dev = dev or self.device
        ref = ReferenceElement(self.dg_orders[0])
        W = ref.mGaussQuadWeights[:, None] * ref.mBasisGQFunVal
        dev.dev_assemble_rhs(0, ref.mGaussQuadNodes, W, terms, self.xin, self.xout,
                             fixes=self.rhs_boundary_fixes(bc_values))

    # ---- upload --------------------------------------------------------------------------------------
    def upload(self, device=0, stream=None, dist=None, options=None):
        """dist = (rank, nranks, nccl_id_bytes) shards the large levels into contiguous element slabs
        (the library decides per level; see amg1d.h).  options: dict for amg1d_set_option, applied
        before the first level (e.g. {"shard_min": 1024})."""
        nL = len(self.levels)
        dev = DeviceHierarchy(nL, device=device, stream=stream, dist=dist)
        for k, v in (options or {}).items():
            dev.set_option(k, v)
        for l, lv in enumerate(self.levels):
            lo, di, up = lv.ops["A"] if lv.explicit else (lv.ops["A"].lo, lv.ops["A"].di, lv.ops["A"].up)
            dinv = blk.to_abi(np.linalg.inv(di))
            if lv.explicit:
                dev.set_level_blocks(l, lo, di, up, dinv, False)
            else:
                dev.set_level_pattern(l, lv.n, lo, di, up, dinv, False, NB, NB)
        for l, (P, ratio) in enumerate(self.transfers):
            dev.set_transfer_pattern(l, self.levels[l].n, P, None, ratio=ratio, period=P.shape[0])
        dev.finalize()
        if dist is not None and dist[1] > 1:
            dev.n_dof[0] = dev.info("local_dofs")          # host vectors are the rank's slab
        self.device = dev
        return dev

    # ---- bookkeeping for the benchmark ---------------------------------------------------
    def dof_updates_per_cycle(self, nPre=3, nPost=3):
        return (nPre + nPost) * sum(lv.n * lv.m for lv in self.levels[:-1])

    def bytes_per_cycle_reference_model(self, nPre=3, nPost=3, with_check=True):
        """B_ref of SURVEY 8d: unfused sweeps on the dense element-block layout."""
        total = 0
        for l, lv in enumerate(self.levels[:-1]):
            n, m = lv.n, lv.m
            mc = self.levels[l + 1].m
            nc = self.levels[l + 1].n
            total += (nPre + nPost) * 8 * n * (4 * m * m + 3 * m)
            total += 8 * (n * (3 * m * m + 2 * m) + n * m * mc + nc * mc)
            total += 8 * (nc * mc + n * m * mc + 2 * n * m)
        if with_check:
            lv = self.levels[0]
            total += 8 * lv.n * (3 * lv.m * lv.m + 2 * lv.m)
        return total

    def tile_rows(self, l):
        """Doubles stored per element on level l: what the device reports for its structure class
        (csrc/layout.cuh) once uploaded, the dense 4 m^2 otherwise."""
        dev = getattr(self, "device", None)
        if dev is not None:
            k = dev.info(f"tile_rows:{l}")
            if k > 0:
                return k
        return 4 * self.levels[l].m ** 2

    def streamed_operator_doubles(self, l):
        """Operator doubles per element a fused leg reads from HBM: tile_rows(l) - minus the m^2 doubles of the
        block-Jacobi inverse where the leg recomputes it (option recompute_dinv) - or 0 when the level's pattern
        table is used (option pattern_resident: the block sets come from a few KB in L1)."""
        dev = getattr(self, "device", None)
        if dev is not None and dev.info("pattern_resident") and dev.info(f"pattern:{l}"):
            return 0
        if dev is not None and dev.info(f"dinv_recompute:{l}") == 1:      # the fused legs invert A_di in registers:
            return self.tile_rows(l) - self.levels[l].m ** 2            # the stored inverse is not read
        return self.tile_rows(l)

    def bytes_per_leg_fused(self, l, down):
        """Algorithmic bytes of one fused leg (f_down / f_up) of level l: the level's operator once
        (tile_rows), b, x in, x out (3 m; the zero-guess down legs of levels > 0 do not read x) and
        the coarse vector (m_c per coarse element); transfer blocks are periodic patterns (L1)."""
        lv, lc = self.levels[l], self.levels[l + 1]
        vec = 2 if (down and l > 0) else 3
        return 8 * (lv.n * (self.streamed_operator_doubles(l) + vec * lv.m) + lc.n * lc.m)

    def bytes_per_cycle_fused(self, with_check=True):
        """Algorithmic bytes of the fused two-kernels-per-level cycle this library runs."""
        return sum(self.bytes_per_leg_fused(l, True) + self.bytes_per_leg_fused(l, False)
                   for l in range(len(self.levels) - 1))


# =====================================================================================================
# CG-first hierarchies on uniform meshes:  CG(p_0) -> ... -> CG(p_k) -> DG / agglomerated levels
# =====================================================================================================
def normalise_transfer(parent, P0, P1, ratio, shift, base):
    """Rewrite explicit transfer blocks for the closed-form parent map
    parent[e] = (e + shift) // ratio + base used by amg1d_set_transfer_pattern.  An element whose only
    parent is the *second* one of the formula gets its block moved from P0 to P1."""
    nf = len(parent)
    pf = (np.arange(nf) + shift) // ratio + base
    Q0 = P0.copy()
    Q1 = np.zeros_like(P0) if P1 is None else P1.copy()
    late = parent == pf + 1
    if np.any((parent != pf) & ~late):
        raise RuntimeError("transfer does not follow the closed-form parent map")
    if np.any(late):
        if np.abs(Q1[late]).max() > 0:
            raise RuntimeError("cannot move a block into an occupied second-parent slot")
        Q1[late] = Q0[late]
        Q0[late] = 0.0
    return Q0, (Q1 if np.abs(Q1).max() > 0 else None)


def compress_transfer(P, period, n_head, n_tail, what=""):
    """(n_fine, mf, mc) explicit blocks -> (n_head + period + n_tail, mf, mc) pattern, checking that
    the interior is periodic."""
    nf = P.shape[0]
    e = np.arange(n_head, nf - n_tail)
    ref = P[n_head + (e - n_head) % period]
    dev = np.abs(P[e] - ref).max() / max(np.abs(P).max(), 1e-300)
    if dev > 1e-11:
        raise RuntimeError(f"{what}: interior transfer blocks are not periodic ({dev:.2e})")
    return np.concatenate([P[:n_head], P[n_head:n_head + period], P[nf - n_tail:]])


def expand_transfer(Ppat, nf, period, n_head, n_tail):
    e = np.arange(nf)
    idx = np.where(e < n_head, e, np.where(e >= nf - n_tail, n_head + period + e - (nf - n_tail),
                                           n_head + (e - n_head) % period))
    return Ppat[idx]


class UniformCgHierarchy:
    """CG(p_0) -> ... -> CG(p_k) -> [DG(q_0) -> ...] -> [agglomerated levels] on a uniform mesh: the
    hierarchy of ``MeshHierarchy(mMeshes, mesh, mBdConds, A; nCG, nDG, nAgg, CDir)``
    (src/mesh_heirarchy.jl:30-138; shapes of tests/cg_heirarchy_test.jl, dg_cg_heirarchy_test.jl,
    full_heirarchy_test.jl and BASELINE C1 / C4) in pattern form.

    CG levels change the polynomial order on a FIXED mesh, so their operators, point-Jacobi diagonals
    and transfers are read off a literal small replica built by the general path (same h, same boundary
    conditions) and compressed to head / interior / tail patterns; the first DG-type level is assembled
    directly on that replica as in the reference, and the levels below it continue with
    ``extend_chain``.  Vectors of CG levels are in *group order*: block k = [vertex k, interior nodes of
    element k] (n + 1 blocks of size p, the last one padded with zeros), see ``blocks.level_slots``.
    """

    HEAD = 8        # explicit head / tail blocks of the CG-side transfers

    def __init__(self, n, cg_orders, dg_orders=(), agg_factors=(), pAgg=1, xin=0.0, xout=1.0,
                 CDir=None, bc_kinds=("neu", "dir")):
        from .agglomerated_dg_mesh import AgglomeratedDgMesh1, uniform_agglomeration
        from .cg_mesh import CgMesh, cg_stiffness
        from .mesh_hierarchy import MeshHierarchy
        self.n = int(n)
        self.cg_orders, self.dg_orders = list(cg_orders), list(dg_orders)
        self.agg_factors, self.pAgg = list(agg_factors), pAgg
        self.xin, self.xout = float(xin), float(xout)
        self.h = (self.xout - self.xin) / self.n
        self.CDir = 1000.0 * n if CDir is None else float(CDir)
        self.bc_kinds = tuple(bc_kinds)
        if not self.cg_orders:
            raise ValueError("At least one CG mesh required.")
        nCG = len(self.cg_orders)
        first_is_dg = bool(self.dg_orders)
        f0 = 1 if first_is_dg or not self.agg_factors else self.agg_factors[0]
        ns = NV * f0
        if self.n < 2 * ns or self.n % f0:
            raise ValueError(f"UniformCgHierarchy needs n >= {2 * ns} (use MeshHierarchy for small meshes)")
        # ---- literal replica by the general path -------------------------------------------------
        mesh = create_uniform_mesh(ns, self.xin, self.xin + ns * self.h)
        bd = set_boundary(mesh, self.xin, self.xin + ns * self.h,
                          [(self.bc_kinds[0], 0.0), (self.bc_kinds[1], 0.0)])
        meshes = [CgMesh(mesh, p) for p in self.cg_orders]
        nDG = nAgg = 0
        if first_is_dg:
            meshes.append(DgMesh(mesh, self.dg_orders[0])); nDG = 1
        elif self.agg_factors:
            meshes.append(AgglomeratedDgMesh1(pAgg, uniform_agglomeration(ns, f0), mesh, meshes[0])); nAgg = 1
        A0 = cg_stiffness(meshes[0], bd)
        Hr = MeshHierarchy(meshes, mesh, [bd] * len(meshes), A0, nCG=nCG, nDG=nDG, nAgg=nAgg,
                           CDir=self.CDir, upload=False)
        self.levels, self.cg_transfers, self.transfers = [], [], []
        slots = [blk.level_slots(m) for m in meshes]
        for l in range(nCG):
            p = self.cg_orders[l]
            lv = UniformLevel(self.n + 1, p)
            lv.explicit = False
            lv.is_cg = True
            lo, di, up = blk.csc_to_blocks(Hr.mStiffness[l], slots[l])
            lv.ops["A"] = compress(lo, di, up, f"CG level {l}")
            self.levels.append(lv)
        # ---- CG-side transfers (two-parent), in the closed-form parent convention ------------------
        for l in range(len(meshes) - 1):
            parent, P0, P1 = blk.transfer_to_blocks(Hr.mInterpolation[l], slots[l], slots[l + 1])
            if l < nCG - 1:
                ratio, shift, base, period = 1, 0, 0, 1            # cg_cg: group k <- groups k, k+1
            elif first_is_dg:
                ratio, shift, base, period = 1, 0, -1, 1           # dg_cg: group k <- elements k-1, k
            else:
                ratio, shift, base, period = f0, f0 - 1, -1, f0    # aggdg_cg: k <- ceil(k/f)-1, +1
            Q0, Q1 = normalise_transfer(parent, P0, P1, ratio, shift, base)
            nh = nt = max(self.HEAD, 2 * period)
            tr = dict(ratio=ratio, shift=shift, base=base, period=period, n_head=nh, n_tail=nt,
                      P0=compress_transfer(Q0, period, nh, nt, f"transfer {l} P0"),
                      P1=None if Q1 is None else compress_transfer(Q1, period, nh, nt, f"transfer {l} P1"))
            self.cg_transfers.append(tr)
        # ---- first DG-type level, then the usual chain -----------------------------------------------
        if len(meshes) > nCG:
            m_d = meshes[nCG].mP + 1
            lv = UniformLevel(self.n // f0, m_d)
            if lv.explicit:
                raise ValueError("mesh too small for the pattern path")
            if first_is_dg:
                Mblock = meshes[nCG].mMassMatrix.mBlocks[NV // 2]
            else:
                Mblock = agg_mass_block(pAgg, f0 * self.h)
            _level_from_window(lv, Hr.mGradient[0], Hr.mDivergence[0], Hr.mC[0], Mblock)
            # the reference assembles this level directly; keep ITS operator, not the recomputed one
            lo, di, up = _blocks(Hr.mStiffness[nCG], NV, m_d)
            lv.ops["A"] = compress(lo, di, up, "first DG-type level")
            self.levels.append(lv)
            if first_is_dg:
                steps = [("dg", q) for q in self.dg_orders[1:]] + [("agg", f) for f in self.agg_factors]
                extend_chain(self.levels, self.transfers, steps, self.n, self.h, self.dg_orders[0], False, pAgg)
            else:
                steps = [("agg", f) for f in self.agg_factors[1:]]
                extend_chain(self.levels, self.transfers, steps, self.n // f0, f0 * self.h, None, True, pAgg)
        self.nCG = nCG
        self._meshes_replica = meshes

    # ---- explicit forms (tests) --------------------------------------------------------------------
    def level_blocks(self, l, name="A"):
        lv = self.levels[l]
        return lv.ops[name] if lv.explicit else lv.ops[name].expand(lv.n)

    def transfer_blocks(self, l):
        """(parent, P0, P1) explicit for transfer l (any kind)."""
        nf = self.levels[l].n
        if l < len(self.cg_transfers):
            t = self.cg_transfers[l]
            parent = (np.arange(nf) + t["shift"]) // t["ratio"] + t["base"]
            ex = lambda P: None if P is None else expand_transfer(P, nf, t["period"], t["n_head"], t["n_tail"])  # noqa: E731
            return parent, ex(t["P0"]), ex(t["P1"])
        P, ratio = self.transfers[l - len(self.cg_transfers)]
        e = np.arange(nf)
        return e // ratio, P[e % ratio], None

    def group_slots(self, l=0):
        """Host DOF (reference numbering, src/cg_mesh.jl:35-45) of every group slot of CG level l."""
        n, p = self.n, self.cg_orders[l]
        s = np.full((n + 1, p), -1, dtype=np.int64)
        s[:, 0] = np.arange(n + 1)
        if p > 1:
            s[:n, 1:] = (n + 1) + np.arange(n)[:, None] * (p - 1) + np.arange(p - 1)[None, :]
        return s

    # ---- right-hand side of cg_stiffness_and_rhs in group order (src/cg_mesh.jl:125-185) -------------
    def rhs(self, func, bc_values, chunk=1 << 20, group_range=None):
        """Right-hand side in group order; ``group_range = (g_begin, g_end)`` returns only that slab of
        groups (multi-GPU: every rank assembles its own slab)."""
        p = self.cg_orders[0]
        ref = ReferenceElement(p)
        n = self.n
        g0, g1 = (0, n + 1) if group_range is None else group_range
        b = np.zeros((g1 - g0, p))

        W = ref.mGaussQuadWeights[:, None] * ref.mBasisGQFunVal            # (nq, p+1)
        for e0 in range(max(g0 - 1, 0), min(g1, n), chunk):                # elements touching the slab
            e1 = min(min(g1, n), e0 + chunk)
            i = np.arange(e0, e1, dtype=np.float64)
            xl = self.xin + (i / n) * (self.xout - self.xin)
            xr = self.xin + ((i + 1) / n) * (self.xout - self.xin)
            hh, xc = xr - xl, (xl + xr) / 2.0
            xq = xc[:, None] + (hh / 2.0)[:, None] * ref.mGaussQuadNodes[None, :]
            fe = (hh / 2.0)[:, None] * (eval_func(func, xq) @ W)           # (chunk, p+1)
            a, z = max(e0, g0), min(e1, g1)                                # own vertex + interior nodes: group e
            if z > a:
                b[a - g0:z - g0, 0] += fe[a - e0:z - e0, 0]
                if p > 1:
                    b[a - g0:z - g0, 1:] = fe[a - e0:z - e0, 2:]
            a, z = max(e0 + 1, g0), min(e1 + 1, g1)                        # right vertex: group e + 1
            if z > a:
                b[a - g0:z - g0, 0] += fe[a - 1 - e0:z - 1 - e0, 1]
        b = b.ravel()
        for slot, val, op in self.rhs_boundary_fixes(bc_values):
            if g0 * p <= slot < g1 * p:
                if op == "set":
                    b[slot - g0 * p] = val
                else:
                    b[slot - g0 * p] += val
        return b

    def rhs_boundary_fixes(self, bc_values):
        """Neumann terms (src/cg_mesh.jl:164-174) and the strong-Dirichlet column terms and rows (:177-182) as an
        ordered list of (global level-0 slot in group order, value, "add" | "set")."""
        p = self.cg_orders[0]
        ref = ReferenceElement(p)
        n = self.n
        fixes = []

        def add(g, col, val):
            fixes.append((g * p + col, float(val), "add"))

        kref = np.einsum("l,li,lj->ij", ref.mGaussQuadWeights, ref.mBasisGQDerivVal, ref.mBasisGQDerivVal)
        for side in (0, 1):                                               # Neumann terms first (:164-174)
            if self.bc_kinds[side] == "neu":
                add(0 if side == 0 else n, 0, (-1.0 if side == 0 else 1.0) * bc_values[side])
        for side in (0, 1):                                               # strong Dirichlet (:177-182)
            if self.bc_kinds[side] != "dir":
                continue
            el = 0 if side == 0 else n - 1
            xl = self.xin + (el / n) * (self.xout - self.xin)
            xr = self.xin + ((el + 1) / n) * (self.xout - self.xin)
            col = (1.0 / ((xr - xl) / 2.0)) * kref[:, side] * bc_values[side]    # A[:, dir] * val on this element
            add(el, 0, -col[0])
            add(el + 1, 0, -col[1])
            for q in range(2, p + 1):
                add(el, q - 1, -col[q])
        for side in (0, 1):
            if self.bc_kinds[side] == "dir":
                fixes.append(((0 if side == 0 else n) * p, float(bc_values[side]), "set"))
        return fixes

    def device_rhs(self, terms, bc_values, dev=None):
        """The right-hand side of cg_stiffness_and_rhs assembled ON THE DEVICE in group order
        (amg1d_dev_assemble_rhs, kind 1); see UniformDgHierarchy.device_rhs."""
        dev = dev or self.device
        ref = ReferenceElement(self.cg_orders[0])
        W = ref.mGaussQuadWeights[:, None] * ref.mBasisGQFunVal
        dev.dev_assemble_rhs(1, ref.mGaussQuadNodes, W, terms, self.xin, self.xout,
                             fixes=self.rhs_boundary_fixes(bc_values))

    # ---- upload ------------------------------------------------------------------------------------------
    def upload(self, device=0, stream=None, dist=None, options=None):
        """dist = (rank, nranks, nccl_id_bytes): contiguous slabs of groups / elements per rank (the
        last rank also holds the closing vertex group of the CG levels).  The two-parent CG transfers
        read ``ratio`` ghost elements more than a fused leg does, hence the deeper default halo."""
        nL = len(self.levels)
        dev = DeviceHierarchy(nL, device=device, stream=stream, dist=dist)
        options = dict(options or {})
        if dist is not None and dist[1] > 1:
            options.setdefault("ghost_depth", 4 + max(t["ratio"] for t in self.cg_transfers))
        for k, v in options.items():
            dev.set_option(k, v)
        for l, lv in enumerate(self.levels):
            if getattr(lv, "is_cg", False):
                pat = lv.ops["A"]
                dinv = np.ascontiguousarray(1.0 / np.einsum("eii->ei", pat.di))       # point Jacobi
                dev.set_level_pattern(l, lv.n, pat.lo, pat.di, pat.up, dinv, True, NB, NB)
            else:
                lo, di, up = lv.ops["A"] if lv.explicit else (lv.ops["A"].lo, lv.ops["A"].di, lv.ops["A"].up)
                dinv = blk.to_abi(np.linalg.inv(di))
                if lv.explicit:
                    dev.set_level_blocks(l, lo, di, up, dinv, False)
                else:
                    dev.set_level_pattern(l, lv.n, lo, di, up, dinv, False, NB, NB)
        for l, t in enumerate(self.cg_transfers):
            dev.set_transfer_pattern(l, self.levels[l].n, t["P0"], t["P1"], ratio=t["ratio"], shift=t["shift"],
                                     base=t["base"], period=t["period"], n_head=t["n_head"], n_tail=t["n_tail"])
        for k, (P, ratio) in enumerate(self.transfers):
            l = len(self.cg_transfers) + k
            dev.set_transfer_pattern(l, self.levels[l].n, P, None, ratio=ratio, period=P.shape[0])
        dev.finalize()
        if dist is not None and dist[1] > 1:
            dev.n_dof[0] = dev.info("local_dofs")          # host vectors are the rank's slab of groups
        self.device = dev
        return dev

    def tile_rows(self, l):
        dev = getattr(self, "device", None)
        if dev is not None:
            k = dev.info(f"tile_rows:{l}")
            if k > 0:
                return k
        lv = self.levels[l]
        return 3 * lv.m ** 2 + (lv.m if getattr(lv, "is_cg", False) else lv.m ** 2)

    def streamed_operator_doubles(self, l):
        """Operator doubles per element a fused leg reads from HBM: tile_rows(l) - minus the m^2 doubles of the
        block-Jacobi inverse where the leg recomputes it (option recompute_dinv) - or 0 when the level's pattern
        table is used (option pattern_resident: the block sets come from a few KB in L1)."""
        dev = getattr(self, "device", None)
        if dev is not None and dev.info("pattern_resident") and dev.info(f"pattern:{l}"):
            return 0
        if dev is not None and dev.info(f"dinv_recompute:{l}") == 1:      # the fused legs invert A_di in registers:
            return self.tile_rows(l) - self.levels[l].m ** 2            # the stored inverse is not read
        return self.tile_rows(l)

    def bytes_per_leg_fused(self, l, down):
        """As UniformDgHierarchy.bytes_per_leg_fused (two-parent transfers read the same coarse
        vector once: neighbouring fine groups share their parents)."""
        lv, lc = self.levels[l], self.levels[l + 1]
        vec = 2 if (down and l > 0) else 3
        return 8 * (lv.n * (self.streamed_operator_doubles(l) + vec * lv.m) + lc.n * lc.m)

    def bytes_per_cycle_fused(self, with_check=True):
        return sum(self.bytes_per_leg_fused(l, True) + self.bytes_per_leg_fused(l, False)
                   for l in range(len(self.levels) - 1))

    def bytes_per_cycle_reference_model(self, nPre=3, nPost=3, with_check=True):
        """B_ref of SURVEY 8d on the dense block layout (point-Jacobi levels: Dinv is m entries)."""
        total = 0
        for l, lv in enumerate(self.levels[:-1]):
            n, m = lv.n, lv.m
            mc, nc = self.levels[l + 1].m, self.levels[l + 1].n
            dv = m if getattr(lv, "is_cg", False) else m * m
            total += (nPre + nPost) * 8 * n * (3 * m * m + dv + 3 * m)
            total += 8 * (n * (3 * m * m + 2 * m) + n * m * mc + nc * mc)
            total += 8 * (nc * mc + n * m * mc + 2 * n * m)
        if with_check:
            lv = self.levels[0]
            total += 8 * lv.n * (3 * lv.m * lv.m + 2 * lv.m)
        return total

    def dof_updates_per_cycle(self, nPre=3, nPost=3):
        tot = 0
        for l, lv in enumerate(self.levels[:-1]):
            tot += (lv.n - 1) * lv.m + 1 if getattr(lv, "is_cg", False) else lv.n * lv.m   # real CG DOFs
        return (nPre + nPost) * tot
