# make_golden.jl - REFERENCE-produced golden vectors for the V-cycle path.
#
# NOT EXECUTED in the build environment (no Julia there; DESIGN.md section 2).  A maintainer runs it once,
# wherever Julia and a checkout of mheinz757/AgglomerationMultigrid1D are available:
#
#     julia julia/make_golden.jl /path/to/AgglomerationMultigrid1D  tests/golden/reference
#
# It drives the UNMODIFIED reference headless - the four tests/*_heirarchy_test.jl shapes and BASELINE's C1
# with exactly their set-up lines, minus `import MATLAB` and the plotting - and writes, per case, raw
# little-endian Float64 / Int64 files that tests/test_reference_golden.py loads when the directory exists:
#
#     <out>/<case>/b.f64  x0.f64  x_after_one_vcycle.f64  x_final.f64  res.f64  err.f64
#     <out>/<case>/iters.txt            (V-cycles to tol 1e-10)
#     <out>/<case>/index_maps.i64       (1-based: per mesh, per element: mNodesInd, then mBaseElementInds and
#                                        mSubAggElementInds of agglomerated elements, in mesh / element order)
#     <out>/<case>/meta.txt             (n, orders, factors, Julia version)
#
# Case names and parameters equal tests/shapes.py (n = 32 so that the literal Python oracle finishes in
# seconds) plus the scripts' own n = 128 for dg_heirarchy.
import LinearAlgebra as la
import SparseArrays as sp

length(ARGS) >= 2 || error("usage: julia make_golden.jl <reference checkout> <output directory>")
const REF = ARGS[1]
const OUT = ARGS[2]

# tests/mesh_generator.jl includes ../src/AgglomerationMultigrid1D.jl and defines create_uniform_mesh, set_boundary!
include(joinpath(REF, "tests", "mesh_generator.jl"))

halving(p, count) = [div(p, 2^(i - 1)) for i in 1:count]           # tempP = div(tempP, 2)

function index_maps(H)
    v = Int64[]
    for m in H.mMeshes, el in m.mElements
        append!(v, el.mNodesInd)
        if el isa aggmg.AbstractAgglomeratedDgElement
            append!(v, el.mBaseElementInds)
            append!(v, el.mSubAggElementInds)
        end
    end
    return v
end

function build(; n, cg = Int64[], dg = Int64[], agg = Int64[], pAgg = 1, xin = 0.0, xout = 1.0)
    CDir = 1000.0 * n
    func(x) = cos(x)
    mesh = create_uniform_mesh(n, xin, xout)
    bdCond = set_boundary!(mesh, xin, xout, [(:neu, -sin(xin)), (:dir, cos(xout))])
    nCG, nDG, nAgg = length(cg), length(dg), length(agg)
    meshes = Vector{aggmg.AbstractMesh}(undef, nCG + nDG + nAgg)
    for (i, p) in enumerate(cg); meshes[i] = aggmg.CgMesh(mesh, p); end
    for (i, p) in enumerate(dg); meshes[nCG + i] = aggmg.DgMesh(mesh, p); end
    base = nCG > 0 ? meshes[1] : (nDG > 0 ? meshes[nDG] : nothing)
    cur = n
    for (i, f) in enumerate(agg)                                    # contiguous ranges, tests/full_heirarchy_test.jl:63-75
        cur = div(cur, f)
        a = [collect((f * (j - 1) + 1):(f * j)) for j in 1:cur]
        meshes[nCG + nDG + i] = i == 1 ? aggmg.AgglomeratedDgMesh1(pAgg, a, mesh, base) :
                                         aggmg.AgglomeratedDgMeshN(pAgg, a, meshes[nCG + nDG + i - 1], base)
    end
    bdConds = Vector{aggmg.BoundaryCondition}(undef, length(meshes))
    fill!(bdConds, bdCond)
    if nCG > 0
        A, b = aggmg.cg_stiffness_and_rhs(meshes[1], mesh, func, bdCond)
        H = aggmg.MeshHierarchy(meshes, mesh, bdConds, A; nCG = nCG, nDG = nDG, nAgg = nAgg, CDir = CDir)
    else
        # the DG-first constructor accepts nAgg but builds no agglomerated levels (src/mesh_heirarchy.jl:140-181):
        # only pure-DG shapes can be produced by the unmodified reference
        nAgg == 0 || error("the reference's DG-first constructor does not build agglomerated levels")
        G, D, C = aggmg.dg_flux_operators(meshes[1], mesh, bdCond, CDir)
        A = C - D * (meshes[1].mMassMatrixLU \ G)
        f, r = aggmg.dg_flux_rhs(meshes[1], mesh, func, bdCond, CDir)
        b = f - D * (meshes[1].mMassMatrixLU \ r)
        H = aggmg.MeshHierarchy(meshes, bdConds, A, G, D, C; nDG = nDG)
    end
    return H, b
end

function run_case(name; kw...)
    H, b = build(; kw...)
    x0 = 0.0 * b
    x1 = aggmg.multigrid_v_cycle(H, x0, b)
    u, iter, res, err = aggmg.multigrid(H, x0, b, 100, 1e-10)
    dir = joinpath(OUT, name)
    mkpath(dir)
    for (fname, v) in (("b", b), ("x0", x0), ("x_after_one_vcycle", x1), ("x_final", u), ("res", res), ("err", err))
        open(joinpath(dir, fname * ".f64"), "w") do io; write(io, htol.(Vector{Float64}(v))); end
    end
    open(joinpath(dir, "index_maps.i64"), "w") do io; write(io, htol.(index_maps(H))); end
    open(joinpath(dir, "iters.txt"), "w") do io; println(io, iter); end
    open(joinpath(dir, "meta.txt"), "w") do io
        println(io, "julia ", VERSION); println(io, kw)
    end
    println(name, ": ", iter, " V-cycles, final residual ", res[end])
end

run_case("cg_heirarchy"; n = 32, cg = halving(8, 4))                                   # tests/cg_heirarchy_test.jl
run_case("dg_heirarchy"; n = 32, dg = halving(8, 4))                                   # tests/dg_heirarchy_test.jl
run_case("dg_heirarchy_n128"; n = 128, dg = halving(8, 4))                             #   at the script's own n
run_case("dg_cg_heirarchy"; n = 32, cg = halving(8, 4), dg = [0])                      # tests/dg_cg_heirarchy_test.jl
run_case("full_heirarchy"; n = 32, cg = halving(8, 4), agg = [4, 2, 2, 2])             # tests/full_heirarchy_test.jl
run_case("C1_cg1_agg"; n = 64, cg = [1], agg = [2])                                    # BASELINE C1, scaled down
run_case("C1_cg1_agg_n1024"; n = 1024, cg = [1], agg = [2])                            # BASELINE C1 at its own size
run_case("C4_cg3_dg1_agg"; n = 32, cg = [3, 1], dg = [1], agg = [2, 2, 2, 2, 2])       # BASELINE C4 shape, scaled down
