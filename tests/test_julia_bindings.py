"""Static check of the Julia wrapper against the C ABI (Julia cannot run in this image): every `ccall((:cs_x, LIB), ret,
(types...), args...)` in clearsky.jl_b200/julia/ClearSkyB200.jl must name an exported symbol, with the arity and the C type
of every argument that include/clearsky_b200.h declares, and every exported cs_* entry point must be bound.  Also walks the
reference's call chain on paper: the methods that `fluxes` (src/fluxes.jl:334), `radiate!` (:377) and `RCM.heating!`
(src/radiative_convective.jl:113) dispatch to exist in the wrapper for the absorber types the reference passes."""
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "clearsky_b200.h")
JULIA = os.path.join(ROOT, "clearsky.jl_b200", "julia", "ClearSkyB200.jl")

# C parameter type -> Julia ccall types that are ABI-compatible with it
SCALARS = {"int32_t": {"Int32", "Cint"}, "int64_t": {"Int64"}, "double": {"Float64", "Cdouble"}}
POINTEES = {"double": "Float64", "int16_t": "Int16", "int32_t": "Int32", "int64_t": "Int64", "uint8_t": "UInt8", "char": "UInt8"}


def c_prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for ret, name, args in re.findall(r"\b(int32_t|const char\s*\*)\s*(cs_\w+)\s*\(([^)]*)\)\s*;", text, flags=re.S):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret.replace(" ", ""), params)
    return protos


def c_param_kind(p):
    """'const double* nu' -> set of acceptable Julia types"""
    p = re.sub(r"\bconst\b", " ", p)
    m = re.match(r"\s*([A-Za-z_0-9]+)\s*((?:\*\s*)*)\s*[A-Za-z_0-9]*\s*$", p)
    assert m, p
    base, stars = m.group(1), m.group(2).count("*")
    if stars == 0:
        return SCALARS[base]
    if base.startswith("cs_") or base == "void":
        inner = "Ptr{Cvoid}"
    else:
        inner = POINTEES[base]
    if stars == 1:
        if inner == "Ptr{Cvoid}":
            return {"Ptr{Cvoid}"}
        ok = {f"Ptr{{{inner}}}", f"Ref{{{inner}}}"}
        if base == "char":
            ok |= {"Cstring", "Ptr{Cchar}"}
        return ok
    assert stars == 2, p
    return {f"Ref{{{inner}}}", f"Ptr{{{inner}}}"} if inner == "Ptr{Cvoid}" else {f"Ref{{Ptr{{{inner}}}}}", f"Ptr{{Ptr{{{inner}}}}}"}


def julia_ccalls():
    src = open(JULIA).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*([\w{}]+),\s*\(([^()]*)\)\s*,?", src, flags=re.S):
        name, ret, types = m.group(1), m.group(2), m.group(3)
        tl = [t.strip() for t in re.split(r",(?![^{}]*\})", types) if t.strip()]
        # count the call arguments that follow the type tuple up to the matching ')'
        depth, i = 1, m.end()
        start = i
        while depth:
            ch = src[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            i += 1
        argtext = src[start:i - 1].strip()
        nargs = 0 if not argtext else len([a for a in re.split(r",(?![^()\[\]{}]*[)\]}])", argtext) if a.strip()])
        out.append((name, ret, tl, nargs, src.count("\n", 0, m.start()) + 1))
    return out


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 60, len(calls)
    for name, ret, types, nargs, line in calls:
        assert name in protos, f"ClearSkyB200.jl:{line}: {name} is not declared in clearsky_b200.h"
        cret, params = protos[name]
        assert ret in ({"Int32", "Cint"} if cret == "int32_t" else {"Cstring", "Ptr{UInt8}", "Ptr{Cchar}"}), (name, ret, line)
        assert len(types) == len(params), f"ClearSkyB200.jl:{line}: {name} binds {len(types)} parameters, the header has {len(params)}"
        assert nargs == len(params), f"ClearSkyB200.jl:{line}: {name} passes {nargs} arguments for {len(params)} parameters"
        for k, (jt, cp) in enumerate(zip(types, params)):
            assert jt in c_param_kind(cp), f"ClearSkyB200.jl:{line}: {name} parameter {k + 1} `{cp}` bound as {jt}"


def test_every_export_is_bound_and_every_binding_exported(cs):
    from clearsky_b200 import _lib
    protos = c_prototypes()
    bound = {c[0] for c in julia_ccalls()}
    assert set(protos) - bound == set(), f"exported but not bound in ClearSkyB200.jl: {sorted(set(protos) - bound)}"
    # the Python twin binds the same set (its SIGNATURES + the two accessor functions), and the .so exports them all
    assert set(_lib.SIGNATURES) | {"cs_last_error", "cs_version"} == set(protos)
    L = _lib.lib()
    for name in protos:
        assert hasattr(L, name), name


def test_reference_call_chain_methods_exist():
    """what the stock callers hand to the core (src/fluxes.jl:334,377; src/radiative_convective.jl:113): ONE unified absorber
    object, a UnifiedAbsorber (src/absorption/absorbers.jl:18-29) or an AcceleratedAbsorber (:114) -- the wrapper must
    define monochromaticfluxes! for B200Discretized on both, and addto! for each member kind they hold"""
    src = open(JULIA).read()
    assert re.search(r"struct\s+B200Discretized\s*<:\s*(ClearSky\.)?AbstractNumericalCore", src)
    sig = re.findall(r"function\s+(?:ClearSky\.)?monochromaticfluxes!\s*\((.*?)\)\s*(?:::Nothing)?\s*(?:where[^\n]*)?\n", src, flags=re.S)
    assert any("B200Discretized" in s for s in sig), "no monochromaticfluxes! method for B200Discretized"
    for T in ("UnifiedAbsorber", "AcceleratedAbsorber", "CIA"):
        assert re.search(rf"addto!\s*\([^)]*::\s*(?:ClearSky\.)?{T}\b", src), f"no addto! method for {T}"
    assert re.search(r"addto!\s*\([^)]*::\s*(?:ClearSky\.)?(?:Abstract)?Gas\b", src), "no addto! method for gases"
    for fn in ("opticaldepth", "radiate!", "jacobian!", "update!"):
        assert re.search(rf"(?:ClearSky\.)?{re.escape(fn)}\s*\(", src), fn
    for sym in ("cs_accel_from_sigma", "cs_sigma_add_accel", "cs_sigma_add_cia", "cs_opticaldepth", "cs_fluxes_batch",
                "cs_table_block", "cs_group_allreduce_sum", "cs_par_read", "cs_rcm_step"):
        # bound AND used by a method (more than its one-line raw binding)
        assert len(re.findall(rf"\b{sym}\b", src)) >= 3, f"{sym} is bound but never called by a wrapper method"
