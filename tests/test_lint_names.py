"""Every test module, GPU-marked ones included, is checked HERE (CPU) for names that are used but never bound:
a GPU test cannot be executed in this container, so a NameError in one would otherwise first show on the B200 box
(round 1: tests/test_vec_ops_gpu.py used `F` without importing it)."""
import ast
import builtins
import glob
import os
import symtable

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FILES = sorted(glob.glob(os.path.join(HERE, "*.py")) + glob.glob(os.path.join(HERE, "hostcheck", "*.py"))
               + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
               + glob.glob(os.path.join(ROOT, "flecsolve_b200", "*.py")) + glob.glob(os.path.join(ROOT, "scripts", "*.py"))
               + glob.glob(os.path.join(ROOT, "scripts", "gpu", "*.py")))


def _module_bindings(tree: ast.Module) -> set[str]:
    """names bound at module level, including inside if/try/with/for blocks of the module body"""
    bound: set[str] = set()

    def visit(body):
        for node in body:
            if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                bound.add(node.name)
            elif isinstance(node, (ast.Import, ast.ImportFrom)):
                for a in node.names:
                    bound.add((a.asname or a.name).split(".")[0])
            elif isinstance(node, (ast.Assign, ast.AugAssign, ast.AnnAssign, ast.For, ast.With, ast.NamedExpr)):
                for t in ast.walk(node):
                    if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Store):
                        bound.add(t.id)
            for attr in ("body", "orelse", "finalbody", "handlers"):
                sub = getattr(node, attr, None)
                if isinstance(sub, list) and not isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                    visit([s for s in sub if isinstance(s, ast.AST)])
            if isinstance(node, ast.ExceptHandler) and node.name:
                bound.add(node.name)
        return bound

    visit(tree.body)
    # `global X` inside a function followed by an assignment also binds a module name
    for node in ast.walk(tree):
        if isinstance(node, ast.Global):
            bound.update(node.names)
    return bound


def undefined_names(path: str) -> list[str]:
    src = open(path).read()
    tree = ast.parse(src, path)
    known = _module_bindings(tree) | set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__"}
    bad: list[str] = []

    def walk(tab: symtable.SymbolTable):
        for sym in tab.get_symbols():
            # a name a scope reads from the global namespace must be bound there
            if sym.is_referenced() and sym.is_global() and not sym.is_assigned() and sym.get_name() not in known:
                bad.append(f"{os.path.relpath(path, ROOT)}: '{sym.get_name()}' in {tab.get_name()}() is never bound")
            if tab.get_type() == "module" and sym.is_referenced() and not (
                    sym.is_assigned() or sym.is_imported() or sym.is_namespace()) and sym.get_name() not in known:
                bad.append(f"{os.path.relpath(path, ROOT)}: '{sym.get_name()}' at module level is never bound")
        for child in tab.get_children():
            walk(child)

    walk(symtable.symtable(src, path, "exec"))
    return bad


@pytest.mark.parametrize("path", FILES, ids=lambda p: os.path.relpath(p, ROOT))
def test_no_unbound_names(path):
    assert undefined_names(path) == []


def test_the_checker_catches_the_round1_bug(tmp_path):
    p = tmp_path / "t.py"
    p.write_text("import numpy as np\n\ndef test_a():\n    from x import y as F\n    F.z()\n\ndef test_b():\n    return F.q(np.zeros(1))\n")
    assert any("'F' in test_b()" in m for m in undefined_names(str(p)))
