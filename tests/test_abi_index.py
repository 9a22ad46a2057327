"""INTEGRATION.md section 6 (the index of the C ABI) is the output of tools/abi_index.py on the current header: every entry point of
include/b200sparse.h is listed with the reference interface it replaces."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_index_matches_header():
    import abi_index

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"<!-- abi-index-begin -->\n(.*?)\n<!-- abi-index-end -->", doc, re.S)
    assert m, "INTEGRATION.md lost its C ABI index"
    assert m.group(1).strip() == abi_index.table().strip(), "run `python tools/abi_index.py` and paste its output into INTEGRATION.md section 6"
    listed = set(re.findall(r"`(b200sp_[a-z0-9_]+)`", m.group(1)))
    assert listed == {s for s, _ in abi_index.entries()}
