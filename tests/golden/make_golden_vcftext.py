#!/usr/bin/env python
"""
Golden VCF text for the writer: the literal `expected` strings of the reference's own CLI tests
(/root/reference/tests/test_transform.py: test_basic, test_basic_maf, test_basic_pgen_input,
test_basic_subset and `ancestry_results` of test_ancestry_from_vcf / test_ancestry_from_bp),
extracted from the test module's source with `ast` -- nothing is retyped.  Each is stored next to
the name of the array fixture (tests/golden/file_*.npz, generated from the same input files by
make_golden.py) whose transform it is the text of.
Build container only:   python tests/golden/make_golden_vcftext.py
"""
from __future__ import annotations

import ast
import json
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference/tests/test_transform.py")

# reference test (or module-level name) -> (array fixture, MAF threshold of the command line)
WANTED = {
    "test_basic": ("file_simple.npz", None),
    "test_basic_maf": ("file_simple.npz", 0.05),
    "test_basic_subset": ("file_simple.npz", None),       # H2/H3 variants removed from the genotypes: H1 only
    "test_basic_pgen_input": ("file_simple.npz", None),   # same cohort read from simple.pgen
    "ancestry_results": ("file_simple_ancestry.npz", None),
}


def main():
    tree = ast.parse(SRC.read_text())
    found = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) in WANTED:
            found[node.targets[0].id] = ast.literal_eval(node.value)
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            for st in node.body:
                if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", None) == "expected":
                    found[node.name] = ast.literal_eval(st.value)
    assert set(found) == set(WANTED), sorted(set(WANTED) - set(found))
    out = {k: {"fixture": WANTED[k][0], "maf": WANTED[k][1], "text": found[k],
               "source": f"tests/test_transform.py::{k}"} for k in WANTED}
    (HERE / "vcftext.json").write_text(json.dumps(out, indent=1))
    for k, v in out.items():
        print(k, len(v["text"].splitlines()), "lines")


if __name__ == "__main__":
    main()
