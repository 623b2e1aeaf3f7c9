"""Static check of the (unexecuted) Julia shim julia/JetsB200.jl against include/jets_b200.h: every
`ccall((:jets_x, LIB), ret, (argtypes...), ...)` must name a declared entry point, pass as many argument
types as the C declaration has parameters, and use a Julia type of the right class for each (pointer vs
integer vs floating point).  Julia is not installed in this image, so this is the strongest check the
shim can get here."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_declarations():
    src = open(os.path.join(ROOT, "include", "jets_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b([A-Za-z_][\w\s\*]*?)\b(jets_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src):
        name, params = m.group(2), m.group(3).strip()
        if params in ("void", ""):
            decls[name] = []
            continue
        kinds = []
        for p in params.split(","):
            p = p.strip()
            if "*" in p or "[" in p or re.match(r"(const\s+)?jets_(buf|op|scalar|dist_op|allgather_fn)\b", p):
                kinds.append("ptr")
            elif re.match(r"(const\s+)?(double|float)\b", p):
                kinds.append("fp")
            else:
                kinds.append("int")
        decls[name] = kinds
    return decls


def split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_kind(t):
    if t.startswith("Ptr{") or t in ("Cstring",) or t.startswith("Ref{"):
        return "ptr"
    if t in ("Cdouble", "Cfloat", "Float64", "Float32"):
        return "fp"
    return "int"


def shim_ccalls():
    src = open(os.path.join(ROOT, "julia", "JetsB200.jl")).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(jets_[a-z0-9_]+),\s*LIB\),\s*([A-Za-z0-9_{}]+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = [t for t in split_top_level(src[m.end():i - 1]) if t]
        calls.append((m.group(1), m.group(2), types, src.count("\n", 0, m.start()) + 1))
    return calls


def test_every_ccall_matches_a_declared_entry_point():
    decls = c_declarations()
    calls = shim_ccalls()
    assert len(decls) >= 90 and len(calls) >= 30
    for name, ret, types, line in calls:
        assert name in decls, f"julia/JetsB200.jl:{line}: {name} is not declared in include/jets_b200.h"
        want = decls[name]
        assert len(types) == len(want), f"julia/JetsB200.jl:{line}: {name} takes {len(want)} arguments, the ccall passes {len(types)}"
        for k, (t, w) in enumerate(zip(types, want)):
            assert julia_kind(t) == w, f"julia/JetsB200.jl:{line}: {name} argument {k + 1}: Julia type {t} vs C parameter class {w}"


def test_nonlinear_leaf_uses_the_linear_view_for_its_jacobian():
    """jets_apply rejects mode DFT on a nonlinear handle, so the pointwise leaf's df!/df'! closures must
    go through a jets_op_as_linear view; and a composition's point! must not allocate host zeros."""
    src = open(os.path.join(ROOT, "julia", "JetsB200.jl")).read()
    body = src[src.index("function JopPointwiseB200"):src.index("function JopStencilB200")]
    assert "unary(:jets_op_as_linear, oh)" in body
    assert "leaf_apply!(m, b200lin, 2, d)" in body and "leaf_apply!(d, b200lin, 1, m)" in body
    assert "leaf_apply!(m, b200, 2, d)" not in body
    assert "function Jets.point!(j::Jets.Jet{D,R,typeof(Jets.JetComposite_f!)}, mₒ::B200Array)" in src


def test_shim_binds_the_hot_path():
    # (jacobian is Jets' own: copy(jet, false) -> deepcopy of the state -> jets_op_clone, then point! ->
    # the leaf's upstate! -> jets_op_set_point)
    bound = {c[0] for c in shim_ccalls()}
    for must in ("jets_init", "jets_buf_create", "jets_buf_upload", "jets_buf_download", "jets_buf_view", "jets_op_diag",
                 "jets_op_pointwise", "jets_op_stencil", "jets_op_dense", "jets_op_compose", "jets_op_sum", "jets_op_block",
                 "jets_op_adjoint", "jets_op_clone", "jets_op_set_point", "jets_apply", "jets_dot", "jets_norm",
                 "jets_lincomb", "jets_op_destroy", "jets_buf_destroy", "jets_last_error",
                 # one call per distributed apply, the host-buffer pipeline, and the solver-loop plumbing
                 "jets_dist_init", "jets_dist_init_host", "jets_dist_op_create", "jets_dist_op_create_dense", "jets_dist_apply",
                 "jets_dist_apply_normal_host", "jets_dist_op_join", "jets_dist_op_destroy", "jets_dist_sum_scalar",
                 "jets_graph_begin", "jets_graph_end", "jets_graph_launch", "jets_apply_axpby", "jets_scalar_create",
                 "jets_scalar_prog", "jets_dot_dev", "jets_norm_dev", "jets_axpby_dev", "jets_op_as_linear"):
        assert must in bound, f"the Julia shim does not bind {must}"
