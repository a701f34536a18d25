"""Host set-up: conversion between the reference's global sparse operators and the element-block
arrays that the C ABI uploads (include/amg1d.h: amg1d_set_level / amg1d_set_transfer).

A level is described by its *slot map* ``slots`` of shape (n_elem, m): slots[e, i] is the host DOF
held by row i of element block e, or -1 for a padding row.
  * DG / agglomerated levels: slots = mesh.mNodesInd (the reference's element-to-DOF map).
  * CG levels: the vertex-first numbering of src/cg_mesh.jl:35-45 is regrouped into n+1 blocks of size
    p, block k = [vertex k, interior nodes of element k]; the last block holds only vertex n.  With
    this grouping the CG stiffness matrix is block tridiagonal.
Blocks are returned in natural (e, i, j) order; ``to_abi`` produces the ABI's column-major layout.
"""
import numpy as np
import scipy.sparse as sp

from .cg_mesh import CgMesh


def level_slots(mesh):
    if isinstance(mesh, CgMesh):
        n, p = mesh.mNodesInd.shape[0], mesh.mP
        slots = np.full((n + 1, p), -1, dtype=np.int64)
        slots[:, 0] = np.arange(n + 1)
        if p > 1:
            slots[:n, 1:] = mesh.mNodesInd[:, 2:]
        return slots
    return np.ascontiguousarray(mesh.mNodesInd, dtype=np.int64)


def slot_lookup(slots, n_dof):
    """elem_of[dof], local_of[dof] for every host DOF."""
    ne, m = slots.shape
    elem_of = np.full(n_dof, -1, dtype=np.int64)
    local_of = np.full(n_dof, -1, dtype=np.int64)
    valid = slots >= 0
    e_idx, i_idx = np.nonzero(valid)
    elem_of[slots[valid]] = e_idx
    local_of[slots[valid]] = i_idx
    if np.any(elem_of < 0):
        raise ValueError("slot map does not cover every DOF of the level")
    return elem_of, local_of


def is_identity_slots(slots):
    return bool(np.array_equal(slots.ravel(), np.arange(slots.size)))


def csc_to_blocks(A, slots, pad_identity=True):
    """Block-tridiagonal blocks (lo, di, up), each (n_elem, m, m) in (e, i, j) order.
    Raises if A has an entry outside the block-tridiagonal band.  Padding rows get a unit diagonal
    (level operators; not the flux operators G, D, C)."""
    A = sp.coo_matrix(A)
    ne, m = slots.shape
    elem_of, local_of = slot_lookup(slots, A.shape[0])
    er, ec = elem_of[A.row], elem_of[A.col]
    d = ec - er
    nz = A.data != 0.0
    if np.any(np.abs(d[nz]) > 1):
        raise ValueError("operator is not block tridiagonal in the given element grouping")
    out = [np.zeros((ne, m, m)) for _ in range(3)]
    for k, off in enumerate((-1, 0, 1)):
        sel = (d == off)
        np.add.at(out[k], (er[sel], local_of[A.row[sel]], local_of[A.col[sel]]), A.data[sel])
    if pad_identity:
        pe, pi = np.nonzero(slots < 0)
        out[1][pe, pi, pi] = 1.0
    return out[0], out[1], out[2]


def blocks_to_csc(lo, di, up, slots, n_dof):
    ne, m = slots.shape
    rows, cols, vals = [], [], []
    for blk, off in ((lo, -1), (di, 0), (up, 1)):
        e = np.arange(ne)
        ok = (e + off >= 0) & (e + off < ne)
        r = np.repeat(slots[e[ok]][:, :, None], m, axis=2)
        c = np.repeat(slots[e[ok] + off][:, None, :], m, axis=1)
        v = blk[ok]
        keep = (r >= 0) & (c >= 0)
        rows.append(r[keep]); cols.append(c[keep]); vals.append(v[keep])
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(n_dof, n_dof))
    A.eliminate_zeros()
    return A


def transfer_to_blocks(L, fine_slots, coarse_slots):
    """parent (n_fine,), P0, P1 (n_fine, m_f, m_c) with
    x_f[e] += P0[e] x_c[parent[e]] + P1[e] x_c[parent[e] + 1].  P1 is None for single-parent
    transfers.  parent may start at -1 (ghost)."""
    L = sp.coo_matrix(L)
    nf, mf = fine_slots.shape
    nc, mc = coarse_slots.shape
    ef, lf = slot_lookup(fine_slots, L.shape[0])
    ecs, lc = slot_lookup(coarse_slots, L.shape[1])
    nz = L.data != 0.0
    fe, ce = ef[L.row[nz]], ecs[L.col[nz]]
    li, lj, val = lf[L.row[nz]], lc[L.col[nz]], L.data[nz]
    pmin = np.full(nf, np.iinfo(np.int64).max, dtype=np.int64)
    pmax = np.full(nf, -1, dtype=np.int64)
    np.minimum.at(pmin, fe, ce)
    np.maximum.at(pmax, fe, ce)
    empty = pmax < 0
    if np.any(empty):          # fine elements with an all-zero row block inherit a neighbour's parent
        pmin[empty] = -2
        for e in np.flatnonzero(empty):
            pmin[e] = pmin[e - 1] if e > 0 and pmin[e - 1] > -2 else -2
        if np.any(pmin == -2):
            first = pmin[pmin > -2][0] if np.any(pmin > -2) else 0
            pmin[pmin == -2] = first
        pmax[empty] = pmin[empty]
    if np.any(pmax - pmin > 1):
        raise ValueError("a fine element depends on more than two adjacent coarse elements")
    two = bool(np.any(pmax > pmin))
    parent = pmin.copy()
    if two:
        # make the map non-decreasing: an element with a single parent q that sits between elements
        # whose first parent is q-1 can equally be written with parent q-1 and its block in P1.
        for e in range(1, nf):
            if parent[e] < parent[e - 1]:
                raise ValueError("transfer parents are not monotone")
    elif np.any(np.diff(parent) < 0):
        raise ValueError("transfer parents are not monotone")
    P0 = np.zeros((nf, mf, mc))
    P1 = np.zeros((nf, mf, mc)) if two else None
    sel0 = ce == parent[fe]
    np.add.at(P0, (fe[sel0], li[sel0], lj[sel0]), val[sel0])
    if two:
        sel1 = ~sel0
        np.add.at(P1, (fe[sel1], li[sel1], lj[sel1]), val[sel1])
    return parent, P0, P1


def to_abi(blk):
    """(e, i, j) -> contiguous column-major-inside-block array as the C ABI expects."""
    return np.ascontiguousarray(np.transpose(blk, (0, 2, 1)), dtype=np.float64)
