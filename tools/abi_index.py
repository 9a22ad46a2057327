"""Index of the C ABI (include/b200sparse.h) for INTEGRATION.md section 6: every entry point with the first sentence of the header
comment that introduces its group, the reference interfaces that comment cites (file:line) and the binding in this repository that
calls it (Kokkos shim header and / or the Python mirror).  `python tools/abi_index.py` prints the markdown table;
tests/test_abi_index.py keeps INTEGRATION.md in step with the header."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def entries():
    src = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    cur, rows = None, []
    for m in re.finditer(r"/\*(.*?)\*/|^(?:int|const char\*|void|int64_t|size_t)\s+(b200sp_[a-z0-9_]+)\s*\(", src, re.S | re.M):
        if m.group(1) is not None:
            txt = " ".join(m.group(1).replace("*", " ").split())
            # a comment that trails a declaration on its line (typedef ...; /* ... */) does not introduce the next group
            line_start = src.rfind("\n", 0, m.start()) + 1
            trailing = src[line_start:m.start()].strip() != ""
            if len(txt) > 40 and not trailing:
                cur = txt
            elif trailing:
                cur = ""
        else:
            rows.append((m.group(2), cur or ""))
    return rows


def bindings(sym):
    out = []
    shim = os.path.join(ROOT, "kokkos-kernels_b200", "kokkos_shim")
    for f in sorted(os.listdir(shim)):
        if re.search(r"\b" + re.escape(sym) + r"\b", open(os.path.join(shim, f)).read()):
            out.append("kokkos_shim/" + f.replace("KokkosSparse_", "").replace("_tpl_spec", ""))
    for f in ("sparse.py", "multigpu.py", "matgen.py"):
        p = os.path.join(ROOT, "kokkos-kernels_b200", f)
        if os.path.exists(p) and re.search(r"\b" + re.escape(sym) + r"\b", open(p).read()):
            out.append(f)
    return out


def table():
    groups = []
    for sym, c in entries():
        if groups and groups[-1][1] == c:
            groups[-1][0].append(sym)
        else:
            groups.append(([sym], c))
    lines = ["| entry points | what (first sentence of the header comment) | reference interface cited there | bound by |", "|---|---|---|---|"]
    for syms, c in groups:
        c = re.sub(r"^-+\s*", "", c)
        c = " ".join(re.sub(r"-{4,}", ". ", c).split())
        first = re.split(r"(?<=[a-z0-9)\]])[.:;] ", c, maxsplit=1)[0].strip(" -")
        first = (first[:157] + "...") if len(first) > 160 else first
        first = first or "status / version of the library"
        refs = [r.rstrip(",-") for r in re.findall(r"[A-Za-z_/]+\.(?:hpp|cpp|h):[0-9][0-9,\-]*", c)]
        b = sorted({x for s in syms for x in bindings(s)})
        lines.append("| " + ", ".join(f"`{s}`" for s in syms) + " | " + first.replace("|", "/") + " | " +
                     (", ".join(f"`{r}`" for r in refs[:3]) or "--") + " | " + (", ".join(b) or "C callers / tests") + " |")
    return "\n".join(lines)


if __name__ == "__main__":
    print(table())
