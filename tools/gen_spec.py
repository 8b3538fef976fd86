#!/usr/bin/env python
"""Spec-driven specialisation: read the reference's OCP description and emit the model / weight header
the kernels are compiled against.

    python tools/gen_spec.py [--ref /root/reference] [--out crazyflie_nmpc_b200/csrc/cf_spec_generated.h] [--check]

Inputs (executed under stub `casadi` / `acados_template` modules backed by sympy -- neither package is needed):
  crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py   state/input symbols and xdot = f(x,u)
  crazyflie_controller/scripts/crazyflie_full_model/generate_c_code.py    N, Tf, W, W_e, yref, bounds, solver choices
Output: one header with
  * CF_SPEC_* constants (sizes, horizon, weights, bounds, default references) and the solver choices as comments
    with a static check that they are the ones the kernels implement (ERK, SQP_RTI, GAUSS_NEWTON, LINEAR_LS, HPIPM);
  * cf_ode_gen / cf_jvp_x_gen / cf_ju_col_gen: f(x,u), the directional derivative (df/dx) d and column j of df/du,
    generated from the symbolic model with common-subexpression elimination.
`--check` regenerates into memory and fails if the committed header differs (tests/test_spec_generated.py).
The kernels keep the sizes nx = 13, nu = 4 (the lane mapping is built around them); everything else about the model
and the cost comes from here.
"""
import argparse
import os
import sys
import types

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class Bag:
    """Attribute bag standing in for AcadosOcp / AcadosModel and their nested option objects."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        b = Bag()
        object.__setattr__(self, name, b)
        return b


class SymVec(list):
    def size(self):
        return (len(self), 1)

    def __sub__(self, other):
        return SymVec(a - b for a, b in zip(self, other))


def stub_modules():
    casadi = types.ModuleType("casadi")

    class SX:
        @staticmethod
        def sym(name, *a):
            return sp.Symbol(name, real=True)

    def vertcat(*items):
        return SymVec(items)

    casadi.SX, casadi.vertcat = SX, vertcat
    casadi.sin, casadi.cos, casadi.tan, casadi.sqrt, casadi.exp = sp.sin, sp.cos, sp.tan, sp.sqrt, sp.exp
    casadi.Function = lambda *a, **k: None
    at = types.ModuleType("acados_template")
    at.AcadosModel = Bag
    at.AcadosOcp = Bag
    at.AcadosOcpSolver = lambda *a, **k: None
    at.AcadosSimSolver = lambda *a, **k: None
    at.AcadosSim = Bag
    return {"casadi": casadi, "acados_template": at}


ALLOWED_IMPORTS = {"casadi", "acados_template", "export_ode_model", "numpy", "scipy", "scipy.linalg", "math"}
FORBIDDEN_NAMES = {"exec", "eval", "compile", "open", "__import__", "input", "globals", "locals", "vars", "getattr", "setattr",
                   "delattr", "breakpoint", "exit", "quit", "os", "sys", "subprocess", "shutil", "socket", "pathlib", "importlib",
                   "ctypes", "pickle", "builtins"}


def vetted_source(path):
    """The two OCP description scripts come from an untrusted tree and are executed to obtain the symbolic model: before
    that, their syntax tree is checked -- imports outside ALLOWED_IMPORTS are dropped, no use of FORBIDDEN_NAMES, no access to
    dunder attributes, no `with` / `try` / class / async / lambda / global constructs.  Anything else aborts the
    generation (the committed csrc/cf_spec_generated.h stays the source of truth)."""
    import ast
    src = open(path).read()
    tree = ast.parse(src, path)
    banned = (ast.With, ast.AsyncWith, ast.Try, ast.ClassDef, ast.AsyncFunctionDef, ast.Lambda, ast.Global, ast.Nonlocal,
              ast.Await, ast.Yield, ast.YieldFrom, ast.Delete, ast.Raise)
    for node in ast.walk(tree):
        bad = None
        if isinstance(node, banned):
            bad = type(node).__name__
        elif isinstance(node, ast.Name) and node.id in FORBIDDEN_NAMES:
            bad = node.id
        elif isinstance(node, ast.Attribute) and node.attr.startswith("__"):
            bad = "." + node.attr
        if bad:
            raise RuntimeError(f"{path}:{getattr(node, 'lineno', 0)}: '{bad}' is not allowed in an OCP description executed by gen_spec")
    # imports outside the allowlist are not executed at all (a later use of their names fails with NameError)
    class DropImports(ast.NodeTransformer):
        def visit_Import(self, node):
            node.names = [a for a in node.names if a.name in ALLOWED_IMPORTS]
            return node if node.names else None

        def visit_ImportFrom(self, node):
            return node if (node.module in ALLOWED_IMPORTS and node.level == 0) else None

    tree = ast.fix_missing_locations(DropImports().visit(tree))
    return compile(tree, os.path.basename(path), "exec")


def load_reference(ref):
    d = os.path.join(ref, "crazyflie_controller", "scripts", "crazyflie_full_model")
    saved = {k: sys.modules.get(k) for k in ("casadi", "acados_template", "export_ode_model")}
    sys.modules.update(stub_modules())
    try:
        em = types.ModuleType("export_ode_model")
        exec(vetted_source(os.path.join(d, "export_ode_model.py")), em.__dict__)
        sys.modules["export_ode_model"] = em
        model = em.export_ode_model()
        g = {"__name__": "generate_c_code", "__file__": os.path.join(d, "generate_c_code.py")}
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            exec(vetted_source(os.path.join(d, "generate_c_code.py")), g)
        return model, g["ocp"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_description(model_py, ocp_py):
    """Generic form of load_reference: a model file defining export_*_ode_model functions (untrusted: vetted) and an OCP
    description that imports it and leaves `ocp` and `model` at module level."""
    mod_name = os.path.splitext(os.path.basename(model_py))[0]
    saved = {k: sys.modules.get(k) for k in ("casadi", "acados_template", mod_name)}
    sys.modules.update(stub_modules())
    try:
        em = types.ModuleType(mod_name)
        exec(vetted_source(model_py), em.__dict__)
        sys.modules[mod_name] = em
        g = {"__name__": "ocp_description", "__file__": ocp_py}
        ALLOWED_IMPORTS.add(mod_name)
        exec(vetted_source(ocp_py), g)
        return g["model"], g["ocp"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


from sympy.printing.c import C99CodePrinter


class Printer(C99CodePrinter):
    """x**2, x**3 as products (no pow calls in device code)."""

    def _print_Pow(self, expr):
        if expr.exp.is_Integer and 2 <= int(expr.exp) <= 3:
            b = self._print(expr.base)
            b = b if expr.base.is_Symbol else f"({b})"
            return "*".join([b] * int(expr.exp))
        return super()._print_Pow(expr)


def ccode(e):
    return Printer().doprint(e)


def c_double(v):
    return repr(float(v))


def c_array(name, vals):
    return f"static constexpr double {name}[{len(vals)}] = {{{', '.join(c_double(v) for v in vals)}}};"


def emit_function(name, args_doc, exprs, out_name, accumulate=False):
    """C body computing out_name[i] (=|+=) exprs[i] with common-subexpression elimination."""
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("t"), optimizations="basic")
    lines = []
    for s, e in repl:
        lines.append(f"    const double {s} = {ccode(e)};")
    for i, e in enumerate(red):
        if accumulate and e == 0:
            continue
        lines.append(f"    {out_name}[{i}] {'+=' if accumulate else '='} {ccode(e)};")
    return lines


MODELS = {
    # name: (model file under the reference tree, OCP description, generated header, source lines of the header comment)
    "crazyflie": (None, None, "cf_spec_generated.h",
                  ["//   crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py  (model)",
                   "//   crazyflie_controller/scripts/crazyflie_full_model/generate_c_code.py   (horizon, cost, bounds, solver choices)"]),
    "pendulum": ("acados/examples/acados_python/pendulum_on_cart/common/pendulum_model.py",
                 os.path.join(ROOT, "tools", "specs", "pendulum_ocp.py"), "cf_spec_pendulum.h",
                 ["//   acados/examples/acados_python/pendulum_on_cart/common/pendulum_model.py  (model)",
                  "//   tools/specs/pendulum_ocp.py  (horizon, cost, bounds of acados/examples/acados_python/tests/test_ocp_setting.py:150-205)"]),
}


def generate(ref, which="crazyflie"):
    model_py, ocp_py, _, src_lines = MODELS[which]
    model, ocp = load_reference(ref) if which == "crazyflie" else load_description(os.path.join(ref, model_py), ocp_py)
    x, u = list(model.x), list(model.u)
    f = list(model.f_expl_expr)
    nx, nu = len(x), len(u)
    assert nx + nu + 1 <= 32, "one row of the stage matrices per lane: nx + nu + 1 <= 32"
    xs = [sp.Symbol(f"x[{i}]", real=True) for i in range(nx)]
    us = [sp.Symbol(f"u[{i}]", real=True) for i in range(nu)]
    ds = [sp.Symbol(f"d[{i}]", real=True) for i in range(nx)]
    sub = dict(zip(x, xs))
    sub.update(zip(u, us))
    fx = [sp.nsimplify(e, rational=False).subs(sub) if False else e.subs(sub) for e in f]
    J = sp.Matrix(fx).jacobian(sp.Matrix(xs))
    Ju = sp.Matrix(fx).jacobian(sp.Matrix(us))
    jvp = list(J * sp.Matrix(ds))
    nnz_x = sum(1 for e in J if e != 0)
    nnz_u = sum(1 for e in Ju if e != 0)
    # leading states that f does not depend on: their column of df/dx is zero, so their forward sensitivities stay [I;0]
    # and their rows of [B';A'] are unit vectors (the feedback program of cf_rti_warp.h neither stores nor multiplies them)
    nfree = 0
    while nfree < nx and all(J[i, nfree] == 0 for i in range(nx)):
        nfree += 1

    W, We = np.asarray(ocp.cost.W, float), np.asarray(ocp.cost.W_e, float)
    assert np.count_nonzero(W - np.diag(np.diag(W))) == 0 and np.count_nonzero(We - np.diag(np.diag(We))) == 0, \
        "the kernels implement diagonal weights"
    Vx, Vu = np.asarray(ocp.cost.Vx, float), np.asarray(ocp.cost.Vu, float)
    assert np.array_equal(Vx, np.vstack([np.eye(nx), np.zeros((nu, nx))])) and \
        np.array_equal(Vu, np.vstack([np.zeros((nx, nu)), np.eye(nu)])), "cost output must be y = [x; u]"
    so = ocp.solver_options
    choices = dict(qp_solver=so.qp_solver, hessian_approx=so.hessian_approx, integrator_type=so.integrator_type,
                   nlp_solver_type=so.nlp_solver_type)
    assert choices == dict(qp_solver="PARTIAL_CONDENSING_HPIPM", hessian_approx="GAUSS_NEWTON", integrator_type="ERK",
                           nlp_solver_type="SQP_RTI"), f"solver choices not implemented by the kernels: {choices}"
    N, Tf = int(ocp.dims.N), float(so.tf)

    out = []
    out.append("// GENERATED by tools/gen_spec.py from the reference's OCP description -- do not edit.")
    out += src_lines
    out.append(f"// solver choices: {', '.join(f'{k} = {v}' for k, v in choices.items())}")
    out.append(f"// structural non-zeros: df/dx {nnz_x}, df/du {nnz_u}")
    out.append("#pragma once")
    out.append(f"#define CF_SPEC_NX {nx}")
    out.append(f"#define CF_SPEC_NU {nu}")
    out.append(f"#define CF_SPEC_N {N}")
    out.append(f"#define CF_SPEC_TF {c_double(Tf)}")
    out.append(f"#define CF_SPEC_NFREE {nfree}   // leading states with a zero column of df/dx (f does not depend on them)")
    out.append("struct CfSpec")
    out.append("{")
    out.append("    " + c_array("W", list(np.diag(W))) + "    // stage weights, cost order y = [x; u]")
    out.append("    " + c_array("W_e", list(np.diag(We))) + "  // terminal weights")
    out.append("    " + c_array("lbu", list(np.asarray(ocp.constraints.lbu, float))))
    out.append("    " + c_array("ubu", list(np.asarray(ocp.constraints.ubu, float))))
    out.append("    " + c_array("yref", list(np.asarray(ocp.cost.yref, float))))
    out.append("    " + c_array("yref_e", list(np.asarray(ocp.cost.yref_e, float))))
    out.append("    " + c_array("x0", list(np.asarray(ocp.constraints.x0, float))))
    out.append("};")
    out.append("")
    out.append("// xdot = f(x,u)")
    out.append("CF_DEV void cf_ode_gen(const double *x, const double *u, double *f)")
    out.append("{")
    out += emit_function("cf_ode_gen", "", fx, "f")
    out.append("}")
    out.append("")
    out.append("// o = (df/dx)(x,u) d")
    out.append("CF_DEV void cf_jvp_x_gen(const double *x, const double *u, const double *d, double *o)")
    out.append("{")
    out.append("    (void) u;")
    out += emit_function("cf_jvp_x_gen", "", jvp, "o")
    out.append("}")
    out.append("")
    out.append("// o += column j of (df/du)(x,u)")
    out.append("CF_DEV void cf_ju_col_gen(const double *x, const double *u, int j, double *o)")
    out.append("{")
    out.append("    (void) x;")
    for i in range(nx):
        cols = [Ju[i, j] for j in range(nu)]
        if all(c == 0 for c in cols):
            continue
        expr = ccode(cols[nu - 1])
        for j in range(nu - 2, -1, -1):
            expr = f"(j == {j}) ? ({ccode(cols[j])}) : ({expr})"
        out.append(f"    o[{i}] += {expr};")
    out.append("}")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--model", default="crazyflie", choices=sorted(MODELS))
    ap.add_argument("--out", default=None)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.out is None:
        a.out = os.path.join(ROOT, "crazyflie_nmpc_b200", "csrc", MODELS[a.model][2])
    text = generate(a.ref, a.model)
    if a.check:
        cur = open(a.out).read() if os.path.exists(a.out) else ""
        if cur != text:
            sys.exit(f"{a.out} is out of date with respect to {a.ref}: run tools/gen_spec.py")
        print("up to date")
        return
    with open(a.out, "w") as fh:
        fh.write(text)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
