"""CPU: the Julia face (krylov.jl_b200/julia/KrylovB200.jl) cannot be executed here (no Julia in the image), so it
is checked statically: every `ccall` is parsed and its symbol, argument count, argument types and return type are
compared with the prototype in include/krylov_b200.h; the three Julia structs that mirror C structs are compared
field by field; and the solver methods must accept the reference's keyword arguments (src/cg.jl:100-117 etc.)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = open(os.path.join(ROOT, "krylov.jl_b200", "julia", "KrylovB200.jl")).read()
HDR = open(os.path.join(ROOT, "include", "krylov_b200.h")).read()


def _strip_comments(c):
    return re.sub(r"/\*.*?\*/", " ", c, flags=re.S)


def _c_prototypes():
    """name -> (return kind, [arg kinds]) with kinds in {'ptr','int','longlong','double','void','cstr'}."""
    txt = _strip_comments(HDR)
    txt = re.sub(r"^\s*#.*$", " ", txt, flags=re.M)          # preprocessor lines
    txt = re.sub(r'extern\s+"C"\s*\{', " ", txt)
    txt = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", txt, flags=re.S)
    txt = re.sub(r"typedef\s+enum\s*\{.*?\}\s*\w+\s*;", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(\w+)\s*\(([^;{}()]*(?:\([^()]*\)[^;{}()]*)*)\)\s*;", txt):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), m.group(3).strip()
        if not (name.startswith("krylov_") or name.startswith("kb200_")):
            continue
        protos[name] = (_kind(ret, is_return=True), [] if args in ("", "void") else [_kind(a) for a in _split_args(args)])
    return protos


def _split_args(args):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    out.append(cur)
    return [a.strip() for a in out]


def _kind(decl, is_return=False):
    d = decl.strip()
    if "*" in d or "(" in d or d.split()[0] in ("KrylovMatvec", "KrylovBlockMatvec"):
        return "cstr" if is_return and "char" in d else "ptr"
    base = re.sub(r"\b(const|unsigned|signed)\b", "", d)
    if is_return:
        base = base.strip()
    else:
        base = " ".join(base.split()[:-1]) if len(base.split()) > 1 else base
    base = base.strip()
    if base == "void":
        return "void"
    if base == "double":
        return "double"
    if base == "long long":
        return "longlong"
    if base in ("int", "KrylovSolverType", "KrylovBlockSolverType", "KrylovDataType", "KrylovDeviceType"):
        return "int"
    if base in ("KrylovWorkspaceOptions", "KrylovOptions", "KrylovB200Options"):
        return "struct"                    # by-value structs (krylov_default_options & co): not called from Julia
    raise AssertionError(f"unhandled C type: {decl!r}")


def _jl_kind(t):
    t = t.strip()
    if t.startswith("Ptr{") or t.startswith("Ref{"):
        return "ptr"
    return {"Cint": "int", "Clonglong": "longlong", "Cdouble": "double", "Cvoid": "void", "Cstring": "cstr"}[t]


def _ccalls():
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*lib\),\s*([\w{}]+),\s*\(", JL):
        i, depth = m.end(), 1
        while depth:                      # matching parenthesis of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(JL[i], 0)
            i += 1
        types = [t for t in _split_args(JL[m.end():i - 1]) if t]
        # the actual arguments: up to the parenthesis closing the ccall
        j, depth, start = i, 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(JL[j], 0)
            j += 1
        actual = [a for a in _split_args(JL[start:j - 1].lstrip(", ")) if a]
        out.append((m.group(1), m.group(2), types, actual))
    return out


def test_every_ccall_matches_the_header():
    protos = _c_prototypes()
    calls = _ccalls()
    assert len(calls) >= 35
    for name, ret, types, actual in calls:
        assert name in protos, f"{name} is not declared in include/krylov_b200.h"
        cret, cargs = protos[name]
        jret = _jl_kind(ret)
        assert jret == cret or (cret == "cstr" and jret in ("cstr", "ptr")), (name, ret, cret)
        assert len(types) == len(cargs), (name, types, cargs)
        assert len(actual) == len(cargs), (name, actual, cargs)
        for t, c in zip(types, cargs):
            jk = _jl_kind(t)
            assert jk == c or (c == "ptr" and jk == "cstr"), (name, t, c)


def _c_struct_fields(name):
    m = re.search(r"typedef\s+struct\s*\{([^{}]*)\}\s*" + name + r"\s*;", _strip_comments(HDR), flags=re.S)
    assert m, name
    out = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        if "(*" in decl:
            out.append(("ptr", re.search(r"\(\*\s*(\w+)\)", decl).group(1)))
            continue
        arr = re.search(r"(\w+)\s*\[(\d+)\]$", decl)
        if arr:
            out.append((f"char[{arr.group(2)}]", arr.group(1)))
            continue
        out.append((_kind(decl), decl.replace("*", " ").split()[-1]))
    return out


def _jl_struct_fields(name):
    m = re.search(r"struct " + name + r"\b.*?\n(.*?)\nend", JL, flags=re.S)
    assert m, name
    body = re.sub(r"#.*", "", m.group(1))
    out = []
    for f in re.split(r"[;\n]", body):
        f = f.strip()
        if not f:
            continue
        fname, ftype = [x.strip() for x in f.split("::")]
        if ftype.startswith("NTuple{"):
            out.append((f"char[{re.search(r'NTuple{(\d+)', ftype).group(1)}]", fname))
        else:
            out.append((_jl_kind(ftype), fname))
    return out


def test_mirrored_structs_have_the_c_layout():
    for jl, c, rename in (("COpts", "KrylovOptions", {}), ("CExt", "KrylovB200Options", {}), ("CStats", "KrylovB200Stats", {})):
        jf, cf = _jl_struct_fields(jl), _c_struct_fields(c)
        assert [k for k, _ in jf] == [k for k, _ in cf], (jl, jf, cf)
        assert [n for _, n in jf] == [n for _, n in cf], (jl, jf, cf)


def test_solver_methods_accept_the_reference_kwargs():
    """src/cg.jl:100-117, minres.jl:138-151, gmres.jl:96-108, bicgstab.jl:105-116: every keyword a user may pass to
    the four solvers is a keyword of fused_solve! (a MethodError otherwise); stats fields are all filled."""
    m = re.search(r"function fused_solve!\(method::Symbol, ws, A::B200CSR\{T\}, b::B200Vector\{T\};(.*?)\) where T", JL, flags=re.S)
    assert m
    kws = set(re.findall(r"(\w+)(?:::[^=]+?)?\s*=(?!=)", m.group(1)))
    for kw in ("M", "N", "ldiv", "radius", "linesearch", "atol", "rtol", "itmax", "timemax", "verbose", "history", "callback",
               "iostream", "λ", "etol", "conlim", "window", "memory", "restart", "reorthogonalization", "c"):
        assert kw in kws, kw
    for field in ("niter", "solved", "inconsistent", "indefinite", "npcCount", "residuals", "Aresiduals", "Acond",
                  "allocation_timer", "timer", "status"):
        assert re.search(r"st\." + field + r"\b|:" + field + r"\b", JL), field
    # documented in INTEGRATION.md: file constructor, adjoint, per-workspace handle (no create/free per solve)
    assert "kb200_csr_read_mtx" in JL and "Base.adjoint(A::B200CSR" in JL and "const HANDLES" in JL
    assert JL.count("(:krylov_workspace_create, lib)") == 1 and "krylov_b200_attach_csr" in JL
    for solver in ("cg!", "minres!", "gmres!", "bicgstab!", "cr!", "cgs!", "cg_lanczos!", "fom!", "fgmres!", "dqgmres!", "diom!"):
        assert solver.rstrip("!") in JL, solver
    # block solver and block-Jacobi preconditioner are bound too
    assert "function Krylov.block_gmres!(" in JL and "(:krylov_block_solve, lib)" in JL and "struct B200Matrix" in JL
    assert "(:krylov_b200_set_preconditioner_blockdiag, lib)" in JL and "struct B200BlockDiagonal" in JL
    m2 = re.search(r"function Krylov\.block_gmres!\(ws::.*?;(.*?)\) where T", JL, flags=re.S)
    bk = set(re.findall(r"(\w+)(?:::[^=]+?)?\s*=(?!=)", m2.group(1)))
    for kw in ("M", "N", "ldiv", "restart", "reorthogonalization", "atol", "rtol", "itmax", "timemax", "verbose", "history", "callback", "iostream"):
        assert kw in bk, kw                      # src/block_gmres.jl:85-97
