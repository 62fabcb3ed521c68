"""Extract every torch_scatter / torch_sparse import and call site of the reference's scripts into
tests/golden/reference_calls.json (run in the build container, where /root/reference exists):

    python tests/golden/make_reference_calls.py [/root/reference]

The fixture pins the drop-in boundary: tests/test_abi.py checks that the shim packages export each
imported name and that each recorded call binds to the shim's signature.  Only names, argument
counts and keyword names are stored — no reference source."""
import ast
import json
import os
import sys

ROOT = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
MODULES = ("torch_scatter", "torch_sparse")
out = {"imports": [], "calls": []}
for d, _, files in os.walk(ROOT):
    for f in sorted(files):
        if not f.endswith(".py"):
            continue
        path = os.path.join(d, f)
        rel = os.path.relpath(path, ROOT)
        try:
            tree = ast.parse(open(path).read())
        except SyntaxError:
            continue
        names = {}  # local name -> (module, attr)
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in MODULES:
                for a in node.names:
                    names[a.asname or a.name] = (node.module, a.name)
                    out["imports"].append({"file": rel, "line": node.lineno, "module": node.module, "name": a.name})
            elif isinstance(node, ast.Import):
                for a in node.names:
                    if a.name.split(".")[0] in MODULES:
                        out["imports"].append({"file": rel, "line": node.lineno, "module": a.name, "name": None})
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in names:
                mod, attr = names[node.func.id]
                out["calls"].append({"file": rel, "line": node.lineno, "module": mod, "name": attr,
                                     "n_positional": len(node.args),
                                     "keywords": [k.arg for k in node.keywords if k.arg]})
out["imports"].sort(key=lambda r: (r["file"], r["line"], r["name"] or ""))
out["calls"].sort(key=lambda r: (r["file"], r["line"]))
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_calls.json")
with open(dst, "w") as fh:
    json.dump(out, fh, indent=1)
print(f"{len(out['imports'])} imports, {len(out['calls'])} call sites -> {dst}")
