"""tests/blockdiagonal_test.jl of the reference, restated for the product's BlockDiagonal / BlockDiagonalLU
(agglomerationmultigrid1d_b200/block_diagonal.py) and, side by side, for the oracle's literal restatement
(oracle/block_diagonal.py).  The script builds three random 3 x 3 blocks, a random sparse 9 x 9 matrix and
prints five difference norms against the dense equivalent (all ~0): one column of A*B through bd_sp_colmul,
A*B, A*b, ALU \\ B, ALU \\ b (tests/blockdiagonal_test.jl:11-44).  Added here: the accumulate-into semantics of
``mul!`` / ``ldiv!`` (src/block_diagonal.jl:172, :305: ``C[inds, :] += ...``), the DimensionMismatch sites
(:167-169, :300-302), scalar products, ``similar`` / ``size`` / ``Matrix`` / ``sparse``, non-contiguous
mBlockInds, and dense' * BlockDiagonal (:181-191)."""
import numpy as np
import pytest
import scipy.sparse as sp

from agglomerationmultigrid1d_b200 import block_diagonal as bd
from oracle import block_diagonal as obd


def _setup(seed=0, nb=3, m=3):
    rng = np.random.default_rng(seed)
    blocks = rng.random((nb, m, m))
    B = sp.random(nb * m, nb * m, density=0.3, random_state=np.random.RandomState(seed), data_rvs=rng.standard_normal,
                  format="csc")
    A2 = np.zeros((nb * m, nb * m))
    for i in range(nb):
        A2[i * m:(i + 1) * m, i * m:(i + 1) * m] = blocks[i]
    return blocks, B, A2


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_blockdiagonal_script(seed):
    blocks, B, A2 = _setup(seed)
    A = bd.BlockDiagonal(blocks)
    Ao = obd.BlockDiagonal([b.copy() for b in blocks])
    Bd = B.toarray()
    # one column through bd_sp_colmul (:13-25): rows of every touched block, values scattered into a dense column
    for col in range(B.shape[1]):
        rows, vals = bd.bd_sp_colmul(A, B, col)
        d = np.zeros(9)
        d[rows] = vals
        assert np.linalg.norm(d - A2 @ Bd[:, col]) <= 1e-14 * 9
        touched = np.unique(B[:, col].nonzero()[0] // 3)
        assert np.array_equal(rows, (touched[:, None] * 3 + np.arange(3)).ravel())     # whole blocks, in order
    # A * B (sparse result, :27-29), product and oracle
    C = A @ B
    assert sp.issparse(C) and np.linalg.norm(C.toarray() - A2 @ Bd) <= 1e-13
    Cs = bd.bd_sp_matmul(A, B)
    assert np.linalg.norm(Cs.toarray() - A2 @ Bd) <= 1e-13
    assert np.linalg.norm((Ao @ B).toarray() - C.toarray()) <= 1e-13
    # A * b (:31-34)
    b = Bd[:, 4 % B.shape[1]]
    assert np.linalg.norm(A @ b - A2 @ b) <= 1e-14 * 9
    # ALU \ B, ALU \ b (:36-44)
    ALU = bd.lu(A)
    F2 = np.linalg.solve(A2, Bd)
    assert np.linalg.norm(ALU.solve(B).toarray() - F2) <= 1e-10 * np.linalg.norm(F2)
    assert np.linalg.norm(bd.bd_sp_solve(ALU, B).toarray() - F2) <= 1e-10 * np.linalg.norm(F2)
    assert np.linalg.norm(Ao.lu().solve(B).toarray() - F2) <= 1e-10 * np.linalg.norm(F2)
    f2 = np.linalg.solve(A2, b)
    assert np.linalg.norm(ALU.solve(b) - f2) <= 1e-10 * np.linalg.norm(f2)
    for col in range(B.shape[1]):
        rows, vals = bd.bd_sp_colsolve(ALU, B, col)
        d = np.zeros(9)
        d[rows] = vals
        assert np.linalg.norm(d - F2[:, col]) <= 1e-10 * max(np.linalg.norm(F2[:, col]), 1.0)


def test_mul_and_ldiv_accumulate_into_their_output():
    """src/block_diagonal.jl:172 and :305 ADD into C: mul!(C, A, B) on a non-zero C gives C0 + A B."""
    blocks, B, A2 = _setup(3)
    A = bd.BlockDiagonal(blocks)
    rng = np.random.default_rng(7)
    for shape in ((9,), (9, 4)):
        X = rng.standard_normal(shape)
        C0 = rng.standard_normal(shape)
        C = C0.copy()
        out = A.mul(C, X)
        assert out is C and np.allclose(C, C0 + A2 @ X, rtol=0, atol=1e-13)
        A.mul(C, X)                                             # a second call adds once more
        assert np.allclose(C, C0 + 2 * (A2 @ X), rtol=0, atol=1e-13)
        C = C0.copy()
        ALU = A.lu()
        out = ALU.ldiv(C, X)
        assert out is C and np.allclose(C, C0 + np.linalg.solve(A2, X), rtol=1e-10, atol=1e-10)
        # `*` and `\` start from zeros (:178-179, :311-312)
        assert np.allclose(A @ X, A2 @ X, atol=1e-13) and np.allclose(ALU.solve(X), np.linalg.solve(A2, X), rtol=1e-10)
    # the oracle's literal loops behave the same way
    Ao = obd.BlockDiagonal([b.copy() for b in blocks])
    X, C0 = rng.standard_normal(9), rng.standard_normal(9)
    C = C0.copy()
    Ao.mul_into(C, X)
    assert np.allclose(C, C0 + A2 @ X, atol=1e-13)


def test_dimension_mismatch_sites():
    blocks, B, A2 = _setup(4)
    A = bd.BlockDiagonal(blocks)
    with pytest.raises(ValueError, match="DimensionMismatch"):
        A.mul(np.zeros(9), np.zeros(8))
    with pytest.raises(ValueError, match="DimensionMismatch"):
        A.mul(np.zeros(8), np.zeros(9))
    with pytest.raises(ValueError, match="DimensionMismatch"):
        A.mul(np.zeros((9, 2)), np.zeros((9, 3)))
    with pytest.raises(ValueError, match="DimensionMismatch"):
        A.lu().ldiv(np.zeros(9), np.zeros(10))
    with pytest.raises(ValueError, match="DimensionMismatch"):
        bd.bd_sp_colmul(A, sp.csc_matrix((8, 2)), 0)
    with pytest.raises(ValueError, match="same size"):
        bd.BlockDiagonal(np.zeros((2, 3, 2)))


def test_utilities_scalar_products_and_general_block_indices():
    blocks, B, A2 = _setup(5)
    A = bd.BlockDiagonal(blocks)
    assert A.size() == (9, 9) and A.size(1) == 9 and A.size(2) == 9 and A.size(3) == 1
    with pytest.raises(ValueError):
        A.size(0)
    S = A.similar()
    assert S.mBlocks.shape == A.mBlocks.shape and S.mBlockSize == 3 and np.array_equal(S.mBlockInds, A.mBlockInds)
    assert np.array_equal(A.toarray(), A2) and np.array_equal(A.tosparse().toarray(), A2)
    assert np.array_equal((A * 2.5).toarray(), 2.5 * A2) and np.array_equal((2.5 * A).toarray(), 2.5 * A2)
    out = A.similar()
    assert A.scale(-1.0, out) is out and np.array_equal(out.toarray(), -A2)
    # dense' * BlockDiagonal (:181-193), accumulating
    rng = np.random.default_rng(9)
    Ad = rng.standard_normal((4, 9))
    assert np.allclose(Ad @ A, Ad @ A2, atol=1e-13)
    C0 = rng.standard_normal((4, 9))
    C = C0.copy()
    A.rmul(C, Ad)
    assert np.allclose(C, C0 + Ad @ A2, atol=1e-13)
    # interleaved (non-contiguous) mBlockInds, as a CG mass matrix would have
    inds = np.array([[0, 3, 6], [1, 4, 7], [2, 5, 8]]).T.copy()        # block k owns DOFs k, k+3, k+6 ... transposed
    inds = np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]]).T[:, [2, 0, 1]].copy()
    G = bd.BlockDiagonal(blocks, 3, inds)
    G2 = np.zeros((9, 9))
    for k in range(3):
        G2[np.ix_(inds[:, k], inds[:, k])] = blocks[k]
    x = rng.standard_normal(9)
    assert np.allclose(G @ x, G2 @ x, atol=1e-13) and np.allclose(G.lu().solve(x), np.linalg.solve(G2, x), rtol=1e-10)
    assert np.array_equal(G.toarray(), G2)
