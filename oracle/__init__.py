"""CPU oracle for the AgglomerationMultigrid1D V-cycle path.  TEST INFRASTRUCTURE ONLY.

This package is a literal, loop-by-loop CPU restatement (numpy / scipy, pure-Python element
loops) of mheinz757/AgglomerationMultigrid1D.  It exists to *check* the CUDA path, never to be
the thing shipped or measured: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
``agglomerationmultigrid1d_b200`` never imports anything from here.

PARITY UNPINNED: the reference stores no golden vectors, known-answer tests or fixtures
(its 18 "tests" are interactive scripts that ``println`` and plot through MATLAB.jl), and the
reference itself cannot be run in this environment (Julia is not installed and there is no
network).  What pins this restatement instead (tests/test_oracle_*.py):

* the ~0 Galerkin-consistency norms the reference scripts print
  (tests/dg_interpolation_test.jl:40-44, tests/aggdg_dg_interpolation_test.jl:46-50,
  tests/aggdg_interpolation_test.jl:59-63, tests/cg_interpolation_test.jl:43);
* exact reproduction of coarse-space polynomials by prolongation;
* discretisation order ~ p+1 (tests/*_convergence_test.jl);
* structural facts (A symmetric positive definite, block tridiagonal, G = -D');
* mesh-independent V-cycle counts (tests/full_heirarchy_test.jl).

Index convention: every index array here is 0-based, i.e. equal to the reference's 1-based
value minus one.  File:line citations are relative to /root/reference.

Modules
    refelem         legendre.jl, gauss_quad.jl, reference_element.jl
    refmesh         meshes.jl, boundary_conditions.jl, tests/mesh_generator.jl
    block_diagonal  block_diagonal.jl
    cg, dg, aggdg   cg_mesh.jl, dg_mesh.jl, agglomerated_dg_mesh.jl
    interpolation   interpolation.jl
    smoother        smoother.jl
    hierarchy       mesh_heirarchy.jl
    solvers         solvers.jl      (the hot path)
    drivers         the tests/*_heirarchy_test.jl script shapes and the BASELINE configs
"""
