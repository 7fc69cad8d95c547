"""Generates tests/golden/golden.{json,npz}: small seeded inputs with the oracle's outputs.

The reference (Rust) cannot be built in this image and ships no golden vectors with inputs for this
path, so these fixtures pin the ORACLE (the reference restatement), after it passed
the reference's own known-answer tests in tests/test_oracle.py.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import KE, clustered, electrolyte, oracle_for, uniform_pm1  # noqa: E402

CASES = [
    dict(name="uniform_2k_containing", gen="uniform_pm1", n=2000, mode=0, theta=0.5, leaf=1, thread=1024),
    dict(name="electrolyte_3k_domain", gen="electrolyte", n=3000, mode=1, theta=1.0, leaf=1, thread=1024),
    dict(name="clustered_3k_leaf8", gen="clustered", n=3000, mode=0, theta=1.0, leaf=8, thread=32),
]

if __name__ == "__main__":
    gens = dict(uniform_pm1=uniform_pm1, electrolyte=electrolyte, clustered=clustered)
    arrays, meta = {}, {"seed": "0xC0FFEE", "k_e": float(KE), "cases": []}
    for case in CASES:
        b = gens[case["gen"]](case["n"])
        o = oracle_for(b, theta=case["theta"], leaf=case["leaf"], thread=case["thread"])
        o.build() if case["mode"] == 0 else o.build_with_domain(b["hw"], b["hh"])
        name = case["name"]
        for k in ("pos", "charge", "radius", "mass", "species"):
            arrays[f"{name}/{k}"] = b[k]
        arrays[f"{name}/perm"] = o.permutation()
        c = o.canonical()
        for f in ("depth", "is_leaf", "start", "end", "path_lo", "charge", "pos"):
            arrays[f"{name}/canon_{f}"] = c[f]
        e, _ = o.field(KE)
        arrays[f"{name}/e_field"] = e
        meta["cases"].append(dict(case, hw=b["hw"], hh=b["hh"]))
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", len(arrays), "arrays")
