"""agglomerationmultigrid1d_b200 - B200-native V-cycle for AgglomerationMultigrid1D.

Host side (this package, Python): the reference's hierarchy construction and solver entry points,
same names and argument meaning (src/AgglomerationMultigrid1D.jl:18-33 include order).  Device side:
``libamg1d.so`` (csrc/, hand-written sm_100a CUDA behind the C ABI of include/amg1d.h).  The solver
entry points have no CPU implementation; without the built library they raise.
"""
from .meshes import Mesh, BoundaryCondition, create_uniform_mesh, set_boundary
from .reference_element import (ReferenceElement, gauss_quad, legendre_val, legendre_val_and_deriv,
                                evaluate_nodal_basis_fun, evaluate_nodal_basis_fun_and_deriv)
from .block_diagonal import BlockDiagonal, BlockDiagonalLU, lu
from .cg_mesh import CgMesh, cg_stiffness, cg_stiffness_and_rhs, cg_rhs
from .dg_mesh import DgMesh, dg_flux_rhs as _dg_flux_rhs
from .agglomerated_dg_mesh import (AgglomeratedDgMesh1, AgglomeratedDgMeshN, agg_dg_flux_rhs,
                                   evaluate_local_modal_basis_fun, evaluate_local_modal_basis_deriv,
                                   uniform_agglomeration)
from .interpolation import (cg_cg_interpolation, dg_dg_interpolation, dg_cg_interpolation,
                            aggdg_aggdg_interpolation, aggdg_dg_interpolation, aggdg_cg_interpolation)
from .smoother import (AbstractSmoother, JacobiSmoother, BlockJacobi, AdditiveSchwarzSmoother,
                       HybridSchwarzSmoother, cg_smoother, dg_smoother)
from .mesh_hierarchy import MeshHierarchy, dg_flux_operators
from .solvers import (multigrid_v_cycle, multigrid, ldiv, pcg, iterative_smoother_solve, apply_smoother)
from .device import DeviceHierarchy
from ._capi import Amg1dError


def dg_flux_rhs(dgMesh, mesh_or_base, func, bdCond, CDir):
    """Dispatch like the reference (src/dg_mesh.jl:342, src/agglomerated_dg_mesh.jl:875)."""
    if isinstance(dgMesh, AgglomeratedDgMesh1):
        return agg_dg_flux_rhs(dgMesh, mesh_or_base, func, bdCond, CDir)
    return _dg_flux_rhs(dgMesh, mesh_or_base, func, bdCond, CDir)
