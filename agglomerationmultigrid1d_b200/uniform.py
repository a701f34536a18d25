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
        n_cur, H = self.n, self.h
        steps = [("dg", p) for p in self.dg_orders[1:]] + [("agg", f) for f in self.agg_factors]
        p_cur, is_agg = p0, False
        for kind, val in steps:
            fine = self.levels[-1]
            if kind == "dg":
                P, ratio, m_c = dg_dg_blocks(val, p_cur), 1, val + 1
                Mblock = (H / 2.0) * ReferenceElement(val).mMassMatrix
                p_cur = val
            else:
                ratio, m_c = val, self.pAgg + 1
                if n_cur % ratio:
                    raise ValueError("agglomeration factor does not divide the element count")
                P = aggdg_aggdg_blocks(self.pAgg, ratio) if is_agg else aggdg_dg_blocks(self.pAgg, p_cur, ratio)
                H *= ratio
                Mblock = agg_mass_block(self.pAgg, H)
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
            self.levels.append(coarse)
            self.transfers.append((P, ratio))
            n_cur = n_next

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
        lv = self.levels[0]
        Dlo, Ddi, Dup = lv.ops["D"] if lv.explicit else lv.ops["D"].expand(min(n, NV))
        Minv = np.linalg.inv(lv.mass)
        e1, e2 = 0, (1 if p0 >= 1 else 0)
        def add(el, vec):                                   # b[el] += vec if el lies in the slab
            if lo_e <= el < hi_e:
                b[(el - lo_e) * m:(el - lo_e + 1) * m] += vec

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
        return b

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

    def bytes_per_cycle_fused(self, with_check=True):
        """Algorithmic bytes of the fused two-kernels-per-level cycle this library runs: per leg the
        level's operator once (4 m^2), b, x in, x out (3 m) and the coarse vector (m_c / ratio);
        transfer blocks are periodic patterns (L1-resident).  Zero-guess down-legs do not read x."""
        total = 0
        for l, lv in enumerate(self.levels[:-1]):
            n, m = lv.n, lv.m
            mc, nc = self.levels[l + 1].m, self.levels[l + 1].n
            down = n * (4 * m * m + (3 if l == 0 else 2) * m) + nc * mc
            up = n * (4 * m * m + 3 * m) + nc * mc
            total += 8 * (down + up)
        return total
