"""Static check of the Python around the library (bench.py, the package, tools/): every name a function loads is bound somewhere
(its own scope, an enclosing function, the module, builtins).  Catches the class of error a GPU-only code path hides from the
CPU suite (a developer bench arm that referred to a variable a refactoring had removed)."""
import ast
import builtins
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted([os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
               + glob.glob(os.path.join(ROOT, "dsd-neo_b200", "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py")))


def _bound(node, skip=None):
    """Names bound directly in this scope (not in nested function scopes, except their own names)."""
    names = set()
    stack = [c for c in ast.iter_child_nodes(node)]
    while stack:
        x = stack.pop()
        if isinstance(x, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(x.name)
            continue  # a nested scope binds its own names
        if isinstance(x, ast.Lambda):
            continue
        if isinstance(x, ast.Name) and isinstance(x.ctx, (ast.Store, ast.Del)):
            names.add(x.id)
        elif isinstance(x, (ast.Import, ast.ImportFrom)):
            names.update((a.asname or a.name).split(".")[0] for a in x.names)
        elif isinstance(x, ast.ExceptHandler) and x.name:
            names.add(x.name)
        elif isinstance(x, (ast.Global, ast.Nonlocal)):
            names.update(x.names)
        stack.extend(ast.iter_child_nodes(x))
    return names


def _args(fn):
    a = fn.args
    names = {x.arg for x in a.args + a.kwonlyargs + getattr(a, "posonlyargs", [])}
    for v in (a.vararg, a.kwarg):
        if v:
            names.add(v.arg)
    return names


def _undefined(tree):
    out = []

    def visit(scope, visible):
        here = visible | _bound(scope)
        if isinstance(scope, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            here |= _args(scope)
        stack = [c for c in ast.iter_child_nodes(scope)]
        while stack:
            x = stack.pop()
            if isinstance(x, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                for d in getattr(x, "decorator_list", []) + x.args.defaults + [k for k in x.args.kw_defaults if k]:
                    stack.append(d)
                visit(x, here)
                continue
            if isinstance(x, ast.ClassDef):
                visit(x, here)
                continue
            if isinstance(x, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                comp = set()
                for g in x.generators:
                    comp |= {n.id for n in ast.walk(g.target) if isinstance(n, ast.Name)}
                inner = here | comp
                for n in ast.walk(x):
                    if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in inner and not hasattr(builtins, n.id):
                        out.append((n.lineno, n.id))
                continue
            if isinstance(x, ast.Name) and isinstance(x.ctx, ast.Load) and x.id not in here and not hasattr(builtins, x.id):
                out.append((x.lineno, x.id))
            stack.extend(ast.iter_child_nodes(x))

    visit(tree, {"__file__", "__name__", "__doc__"})
    return sorted(set(out))


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(p, ROOT) for p in FILES])
def test_every_loaded_name_is_bound(path):
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    assert _undefined(tree) == []
