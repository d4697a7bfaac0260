#!/usr/bin/env python
"""Mechanical API comparison with the reference tree (read-only): for every module under ``<ref>/src/neuronx_distributed``
that also exists here, report public functions / classes / methods / module constants that are missing, and parameters of the
reference's signatures that this package's callable does not accept.  Used to drive docs/MIGRATION.md "Call-compatibility".

    python tools/api_sweep.py [/root/reference/src/neuronx_distributed]
"""
from __future__ import annotations

import ast
import importlib
import inspect
import os
import sys

PKG = "neuronx_distributed_b200"


def _params(fn: ast.FunctionDef):
    a = fn.args
    return [x.arg for x in a.posonlyargs + a.args + a.kwonlyargs if x.arg not in ("self", "cls")]


def sweep(ref: str):
    missing_modules, missing_names, missing_params = [], [], []
    for root, _, files in os.walk(ref):
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), ref)
            mod = rel[:-3].replace("/", ".")
            if mod.endswith("__init__"):
                mod = mod[:-9].rstrip(".")
            try:
                tree = ast.parse(open(os.path.join(root, f)).read())
            except SyntaxError:
                continue
            public = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and not n.name.startswith("_")]
            if not public:
                continue
            try:
                m = importlib.import_module(PKG + ("." + mod if mod else ""))
            except ImportError as e:
                missing_modules.append((mod, str(e)[:80]))
                continue
            for n in tree.body:
                if isinstance(n, ast.Assign):
                    for t in n.targets:
                        if isinstance(t, ast.Name) and not t.id.startswith("_") and t.id != "logger" and not hasattr(m, t.id):
                            missing_names.append((mod, t.id))
            for n in public:
                obj = getattr(m, n.name, None)
                if obj is None:
                    missing_names.append((mod, n.name))
                    continue
                targets = [(n.name, n, obj)] if isinstance(n, ast.FunctionDef) else [
                    (f"{n.name}.{b.name}", b, getattr(obj, b.name, None)) for b in n.body
                    if isinstance(b, ast.FunctionDef) and (b.name == "__init__" or not b.name.startswith("_"))]
                for name, node, target in targets:
                    if target is None:
                        missing_names.append((mod, name))
                        continue
                    try:
                        sig = inspect.signature(target)
                    except (TypeError, ValueError):
                        continue
                    if any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values()):
                        continue
                    lacking = [w for w in _params(node) if w not in sig.parameters]
                    if lacking:
                        missing_params.append((mod, name, lacking))
    return missing_modules, missing_names, missing_params


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/neuronx_distributed"
    mods, names, params = sweep(ref_root)
    print(f"modules without a counterpart: {len(mods)}")
    for r in sorted(mods):
        print("  ", r)
    print(f"public names without a counterpart: {len(names)}")
    for r in sorted(names):
        print("  ", r)
    print(f"callables that do not accept a reference parameter name: {len(params)}")
    for r in sorted(params):
        print("  ", r)
