"""Host set-up parity: the product's vectorised assembly against the oracle's literal restatement.

Index maps (element-to-DOF, agglomeration maps, block index tables) must be bit-exact; floating-point
operators agree to rounding.  Also covers the element-block conversion that feeds the C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

import agglomerationmultigrid1d_b200 as aggmg
from agglomerationmultigrid1d_b200 import blocks as blk
from shapes import SHAPES, build_oracle, build_package


@pytest.fixture(scope="module", params=sorted(SHAPES))
def pair(request):
    kw = SHAPES[request.param]
    Ho, x0, bo, _ = build_oracle(**kw)
    Hp, _, bp = build_package(upload=False, **kw)
    return request.param, Ho, bo, Hp, bp


def test_index_maps_bit_exact(pair):
    name, Ho, bo, Hp, bp = pair
    for mo, mp in zip(Ho.mMeshes, Hp.mMeshes):
        nodes_o = np.stack([np.asarray(el.mNodesInd) for el in mo.mElements])
        assert np.array_equal(nodes_o, mp.mNodesInd)
        assert mo.mNumNodes == mp.mNumNodes
        if hasattr(mo.mElements[0], "mBaseElementInds"):
            for k, el in enumerate(mo.mElements):
                ep = mp.mElements[k]
                # oracle keeps the reference's 1-based element ids
                assert np.array_equal(np.asarray(el.mBaseElementInds) - 1, ep.mBaseElementInds)
                assert np.array_equal(np.asarray(el.mSubAggElementInds) - 1, ep.mSubAggElementInds)
    for So, Sp in zip(Ho.mSmoothers, Hp.mSmoothers):
        if hasattr(So, "mBlockInds"):
            assert np.array_equal(So.mBlockInds, Sp.mBlockInds)


def test_operators_agree(pair):
    name, Ho, bo, Hp, bp = pair
    assert np.abs(bo - bp).max() <= 1e-13 * np.abs(bo).max()
    for l, (So, Sp) in enumerate(zip(Ho.mStiffness, Hp.mStiffness)):
        assert So.shape == Sp.shape
        assert abs(So - Sp).max() <= 1e-13 * abs(So).max(), (name, l)
    for l, (Io, Ip) in enumerate(zip(Ho.mInterpolation, Hp.mInterpolation)):
        assert abs(sp.csc_matrix(Io) - sp.csc_matrix(Ip)).max() <= 1e-13, (name, l)


def test_block_conversion_roundtrip(pair):
    """csc -> (lo, di, up) element blocks -> csc is lossless, CG permutation included; the transfer
    blocks reproduce L entry by entry."""
    name, Ho, bo, Hp, bp = pair
    slots = [blk.level_slots(m) for m in Hp.mMeshes]
    for l, A in enumerate(Hp.mStiffness):
        lo, di, up = blk.csc_to_blocks(A, slots[l])
        assert np.all(lo[0] == 0.0) and np.all(up[-1] == 0.0)
        back = blk.blocks_to_csc(lo, di, up, slots[l], A.shape[0])
        assert abs(back - A).max() == 0.0
    for l, L in enumerate(Hp.mInterpolation):
        parent, P0, P1 = blk.transfer_to_blocks(L, slots[l], slots[l + 1])
        assert np.all(np.diff(parent) >= 0)
        nf, mf = slots[l].shape
        nc, mc = slots[l + 1].shape
        rng = np.random.default_rng(l)
        xc_host = rng.standard_normal(L.shape[1])
        xc = np.zeros((nc + 2, mc))          # ghost element on both sides
        valid = slots[l + 1] >= 0
        xc[1:-1][valid] = xc_host[slots[l + 1][valid]]
        xf = np.einsum("eij,ej->ei", P0, xc[parent + 1])
        if P1 is not None:
            xf += np.einsum("eij,ej->ei", P1, xc[parent + 2])
        ref = sp.csc_matrix(L) @ xc_host
        vf = slots[l] >= 0
        assert np.abs(xf[vf] - ref[slots[l][vf]]).max() <= 1e-13 * max(1.0, np.abs(ref).max())


def test_reference_argument_errors():
    mesh = aggmg.create_uniform_mesh(8, 0.0, 1.0)
    bd = aggmg.set_boundary(mesh, 0.0, 1.0, [("neu", 0.0), ("dir", 1.0)])
    m = aggmg.CgMesh(mesh, 2)
    A, b = aggmg.cg_stiffness_and_rhs(m, mesh, np.cos, bd)
    with pytest.raises(ValueError, match="At least one CG mesh"):
        aggmg.MeshHierarchy([m], mesh, [bd], A, nCG=0, upload=False)
    with pytest.raises(ValueError, match="does not match"):
        aggmg.MeshHierarchy([m], mesh, [bd], A, nCG=2, upload=False)
    dgm = aggmg.DgMesh(mesh, 1)
    with pytest.raises(ValueError, match="p = 0 and p = 1"):
        aggmg.AgglomeratedDgMesh1(2, [[0, 1], [2, 3], [4, 5], [6, 7]], mesh, dgm)
    with pytest.raises(ValueError):
        aggmg.AgglomeratedDgMesh1(1, [[0, 2], [1, 3], [4, 5], [6, 7]], mesh, dgm)   # not contiguous
    with pytest.raises(ValueError, match="interpFlag"):
        aggmg.dg_cg_interpolation(dgm, m, mesh, 3)


def test_empty_and_edge_meshes():
    """n = 1 (both vertices are boundary vertices) and a single-level hierarchy."""
    Ho, x0, bo, _ = build_oracle(1, dg_orders=[2])
    Hp, _, bp = build_package(1, dg_orders=[2], upload=False)
    assert abs(Ho.mStiffness[0] - Hp.mStiffness[0]).max() <= 1e-13 * abs(Ho.mStiffness[0]).max()
    assert np.abs(bo - bp).max() <= 1e-13 * np.abs(bo).max()
    with pytest.raises(ValueError):
        aggmg.create_uniform_mesh(0, 0.0, 1.0)
