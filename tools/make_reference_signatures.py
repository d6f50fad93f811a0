"""Record the keyword signatures of the reference's public entry points on the hot path.

Run in the builder container (needs /root/reference; parses the source with `ast`, JAX is not needed):
    python tools/make_reference_signatures.py
Writes tests/golden/reference_signatures.json, which tests/test_host_and_abi.py checks the host mirror
against on any box (the GPU box has no /root/reference).
Reference: src/discoeb/perturbations.py:926-997, 1000-1061, 1065-1224.
"""
import ast
import json
import os
import warnings

REF = "/root/reference/src/discoeb/perturbations.py"
NAMES = ["evolve_perturbations", "evolve_perturbations_batched", "get_power", "get_power_smoothed",
         "power_Kaiser", "power_multipoles", "get_xi_from_P"]


def signatures(path, names):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            a = node.args
            out[node.name] = {
                "lineno": node.lineno,
                "positional": [x.arg for x in a.args],
                "keyword_only": [[x.arg, None if d is None else ast.unparse(d)]
                                 for x, d in zip(a.kwonlyargs, a.kw_defaults)],
            }
    return out


if __name__ == "__main__":
    sig = signatures(REF, NAMES)
    missing = [n for n in NAMES if n not in sig]
    assert not missing, missing
    dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_signatures.json")
    with open(dst, "w") as f:
        json.dump(sig, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(dst), len(sig), "functions")
