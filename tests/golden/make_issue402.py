"""Extract the reference's regression fixture `sparse/unit_test/matrixIssue402.hpp`
(1813 x 1813, 11156 nnz circuit matrix, Test_Sparse_spgemm.hpp:372-442) into
tests/golden/issue402.npz.  Run in the build container (needs /root/reference):
    python tests/golden/make_issue402.py
The fixture is DATA the reference's own test feeds to spgemm; no code is copied."""
import os
import re

import numpy as np

SRC = "/root/reference/sparse/unit_test/matrixIssue402.hpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "issue402.npz")


def grab(text, name):
    m = re.search(name + r"\s*\[\s*\d+\s*\]\s*=\s*\{(.*?)\};", text, re.S)
    return m.group(1).replace("\n", " ")


def main():
    t = open(SRC).read()
    values = np.array([float(x) for x in grab(t, "values").split(",") if x.strip()], dtype=np.float64)
    rowmap = np.array([int(x) for x in grab(t, "rowmap").split(",") if x.strip()], dtype=np.int32)
    entries = np.array([int(x) for x in grab(t, "entries").split(",") if x.strip()], dtype=np.int32)
    assert len(rowmap) == 1814 and len(entries) == 11156 and len(values) == 11156, (len(rowmap), len(entries), len(values))
    np.savez_compressed(OUT, rowmap=rowmap, entries=entries, values=values)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
