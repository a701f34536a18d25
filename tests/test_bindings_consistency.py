"""The three bindings of the C ABI agree on every prototype, argument by argument:
include/amg1d.h (the contract), agglomerationmultigrid1d_b200/_capi.py (ctypes, the tested path) and
julia/device_hierarchy.jl (the reference-side ccall binding, which cannot be executed here - so its
signatures are checked textually), plus the Julia excerpt shown in INTEGRATION.md."""
import ctypes as C
import os
import re

from agglomerationmultigrid1d_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_TYPES = {
    "amg1d_t*": "ptr", "const amg1d_t*": "ptr", "amg1d_t**": "ptr*", "void*": "ptr", "const void*": "ptr",
    "void**": "ptr*", "int": "i32", "int64_t": "i64", "uint64_t": "u64", "double": "f64",
    "double*": "f64*", "const double*": "f64*", "int64_t*": "i64*", "const int64_t*": "i64*",
    "int*": "i32*", "const int*": "i32*", "const char*": "str", "double*const*": "f64**",
    "const double*const*": "f64**",
}
CT_TYPES = {
    C.c_void_p: "ptr", C.POINTER(C.c_void_p): "ptr*", C.c_int: "i32", C.c_int64: "i64", C.c_uint64: "u64",
    C.c_double: "f64", C.POINTER(C.c_double): "f64*", C.POINTER(C.c_int64): "i64*",
    C.POINTER(C.c_int): "i32*", C.c_char_p: "str", C.POINTER(C.POINTER(C.c_double)): "f64**",
}
JL_TYPES = {
    "Ptr{Cvoid}": "ptr", "Ref{Ptr{Cvoid}}": "ptr*", "Cint": "i32", "Int64": "i64", "UInt64": "u64",
    "Float64": "f64", "Ptr{Float64}": "f64*", "Ptr{Int64}": "i64*", "Ref{Cint}": "i32*", "Ptr{Cint}": "i32*",
    "Cstring": "str", "Ptr{Ptr{Float64}}": "f64**",
}


def header_prototypes():
    text = open(os.path.join(ROOT, "include", "amg1d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w \*]*?)\s*\b(amg1d_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        kinds = []
        for a in [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]:
            a = re.sub(r"\s+", " ", a)
            m = re.match(r"(.*?[\s\*])([A-Za-z_]\w*)$", a)           # type, then the parameter name
            typ = (m.group(1) if m else a).replace(" *", "*").replace("* ", "*").strip()
            kinds.append(C_TYPES[typ])
        ret = re.sub(r"\s+", " ", ret.strip()).replace(" *", "*")
        out[name] = (C_TYPES.get(ret, ret), kinds)
    return out


def test_header_parser_sees_every_symbol():
    protos = header_prototypes()
    assert sorted(protos) == sorted(capi.PROTOTYPES)


def test_ctypes_table_matches_header_argument_by_argument():
    protos = header_prototypes()
    for name, (res, args) in capi.PROTOTYPES.items():
        hret, hargs = protos[name]
        assert [CT_TYPES[a] for a in args] == hargs, name
        assert CT_TYPES[res] == hret or (hret == "str" and res is C.c_char_p), (name, hret)


def julia_ccalls(text):
    calls = []
    for m in re.finditer(r"ccall\(\(:(amg1d_\w+), libamg1d\),\s*(\w+),\s*\(([^)]*)\)", text):
        types = [t.strip() for t in m.group(3).replace("\n", " ").split(",") if t.strip()]
        calls.append((m.group(1), m.group(2), types))
    return calls


def test_julia_binding_matches_header():
    protos = header_prototypes()
    text = open(os.path.join(ROOT, "julia", "device_hierarchy.jl")).read()
    calls = julia_ccalls(text)
    assert len(calls) >= 19
    for name, ret, types in calls:
        hret, hargs = protos[name]                                   # KeyError: binding calls an unknown symbol
        assert [JL_TYPES[t] for t in types] == hargs, (name, types, hargs)
        assert JL_TYPES[ret] == hret, (name, ret, hret)
    # the entry points of the hot path and its boundary are all bound
    bound = {c[0] for c in calls}
    for need in ("amg1d_create", "amg1d_set_level", "amg1d_set_transfer", "amg1d_finalize", "amg1d_destroy",
                 "amg1d_vcycle", "amg1d_solve", "amg1d_ldiv", "amg1d_pcg", "amg1d_apply_smoother",
                 "amg1d_smoother_solve", "amg1d_matvec", "amg1d_residual", "amg1d_restrict", "amg1d_prolong",
                 "amg1d_direct_solve", "amg1d_set_option", "amg1d_get_info", "amg1d_last_error"):
        assert need in bound, need
    # the drop-in methods carry the reference's own signatures (src/solvers.jl:19-20, :63, :84, :116-117)
    for sig in ("multigrid_v_cycle(H::MeshHierarchy, x0::AbstractVector, b::AbstractVector;",
                "multigrid(H::MeshHierarchy, x0::AbstractVector, b::AbstractVector, maxiter::Integer, tol::AbstractFloat)",
                "ldiv!(y::AbstractVector, H::MeshHierarchy, b::AbstractVector)",
                "ldiv!(H::MeshHierarchy, b::AbstractVector)"):
        assert sig in text, sig


def test_integration_md_excerpt_is_the_julia_file():
    """Every ccall shown in INTEGRATION.md is, verbatim, a ccall of julia/device_hierarchy.jl."""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    jl = open(os.path.join(ROOT, "julia", "device_hierarchy.jl")).read()
    shown = julia_ccalls(md)
    assert len(shown) >= 8
    have = julia_ccalls(jl)
    for call in shown:
        assert call in have, call
