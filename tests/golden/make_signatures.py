"""Freeze the constructor / forward signatures of the reference's CausalGCN / CausalGAT / CausalGIN (model.py) into
tests/golden/signatures.json -- read with `ast`, the module is never imported.  Run in the build container only:

    python tests/golden/make_signatures.py
"""
import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CAL_REFERENCE", "/root/reference")


def reference_signatures(ref=REF):
    tree = ast.parse(open(os.path.join(ref, "model.py")).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("CausalGCN", "CausalGAT", "CausalGIN"):
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in ("__init__", "forward"):
                    out["%s.%s" % (node.name, f.name)] = {
                        "args": [x.arg for x in f.args.args],
                        "defaults": [ast.literal_eval(d) for d in f.args.defaults],
                        "line": f.lineno}
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "signatures.json"), "w") as fh:
        json.dump(reference_signatures(), fh, indent=1, sort_keys=True)
        fh.write("\n")
