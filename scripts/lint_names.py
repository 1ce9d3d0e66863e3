"""Minimal undefined-name check (no linter is installed in the build image): every name a function loads must be bound
somewhere in that function, an enclosing function, the module, or builtins.  Flow-insensitive on purpose -- it finds
typos and stale references after refactors, which is what an untested code path needs.

    python scripts/lint_names.py bench.py cvd_b200/*.py
"""
import ast
import builtins
import sys


def bound_names(node):
    """Names bound directly in the scope of ``node`` (a Module, FunctionDef or Lambda), nested scopes excluded."""
    names = set()
    if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
        a = node.args
        for arg in a.posonlyargs + a.args + a.kwonlyargs:
            names.add(arg.arg)
        if a.vararg:
            names.add(a.vararg.arg)
        if a.kwarg:
            names.add(a.kwarg.arg)

    def visit(n):
        for child in ast.iter_child_nodes(n):
            if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                names.add(child.name)
                continue                      # separate scope
            if isinstance(child, ast.Lambda):
                continue
            if isinstance(child, (ast.Import, ast.ImportFrom)):
                for al in child.names:
                    names.add((al.asname or al.name).split(".")[0])
            if isinstance(child, ast.Name) and isinstance(child.ctx, (ast.Store, ast.Del)):
                names.add(child.id)
            if isinstance(child, ast.ExceptHandler) and child.name:
                names.add(child.name)
            if isinstance(child, (ast.Global, ast.Nonlocal)):
                names.update(child.names)
            visit(child)
    visit(node if not isinstance(node, ast.Lambda) else ast.Expr(value=node.body))
    return names


def check(path):
    tree = ast.parse(open(path).read(), path)
    problems = []
    module_names = bound_names(tree) | set(dir(builtins)) | {"__file__", "__name__", "__doc__"}

    def walk(node, chain):
        scope = chain + [bound_names(node)] if not isinstance(node, ast.Module) else chain

        def visit(n):
            for child in ast.iter_child_nodes(n):
                if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                    for d in getattr(child, "decorator_list", []):
                        # comprehension targets inside a decorator expression are bound there
                        local = {n.id for n in ast.walk(d) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store)}
                        scope.append(local)
                        visit(d)
                        scope.pop()
                    for dflt in child.args.defaults + [d for d in child.args.kw_defaults if d is not None]:
                        visit(ast.Expr(value=dflt))
                    walk(child, scope)
                    continue
                if isinstance(child, ast.ClassDef):
                    walk_class(child, scope)
                    continue
                if isinstance(child, ast.Name) and isinstance(child.ctx, ast.Load):
                    if not any(child.id in s for s in scope) and child.id not in module_names:
                        problems.append((path, child.lineno, child.id))
                visit(child)
        body = node if not isinstance(node, ast.Lambda) else ast.Expr(value=node.body)
        visit(body)

    def walk_class(node, chain):
        cls_names = bound_names(node)
        for child in node.body:
            if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)):
                walk(child, chain)              # class scope is not visible from methods
            else:
                for sub in ast.walk(child):
                    if isinstance(sub, ast.Name) and isinstance(sub.ctx, ast.Load):
                        if sub.id not in cls_names and not any(sub.id in s for s in chain) and sub.id not in module_names:
                            problems.append((path, sub.lineno, sub.id))

    walk(tree, [])
    return problems


if __name__ == "__main__":
    bad = []
    for p in sys.argv[1:]:
        bad += check(p)
    for path, line, name in bad:
        print(f"{path}:{line}: undefined name {name!r}")
    sys.exit(1 if bad else 0)
