#!/usr/bin/env python
"""Dynamic cost per SOURCE LINE of one profiled kernel: joins the SASS rows of an .ncu-rep (executed instructions,
stall samples, shared-memory wavefronts per instruction) with the line table of the in-tree object file (nvdisasm -g).

  python tools/ncu_lines.py gpurun_out/r02c6_esc.ncu-rep spgemm esc_num_kernel [top]

`unit` names kokkos-kernels_b200/lib/<unit>.cu.o; the kernel is the first one in the report whose name contains the
given substring, and the cubin function with the same mangled template arguments.  The object file must be the build
the report was taken from (instruction counts are compared)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("B200SP_OBJ_DIR", os.path.join(ROOT, "kokkos-kernels_b200", "lib"))  # B200SP_OBJ_DIR: objects of the profiled commit
CSRC_DIR = os.environ.get("B200SP_SRC_DIR", os.path.join(ROOT, "kokkos-kernels_b200", "csrc"))


def ncu_rows(rep, needle):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + needle], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the first kernel block only
    name = rows[0][1]
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr):
            data.append(r)
    return name, hdr, data


def disasm_lines(unit, demangled):
    obj = os.path.join(LIB, unit + ".cu.o")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # function whose demangled name matches
    funcs = re.findall(r"^\s*\.global\s+(\S+)", text, flags=re.M)
    want = re.sub(r"\(int\)", "", demangled)
    want = re.sub(r"\s+", "", want.split("(")[0].replace("void ", "").replace("b200sp::", ""))
    target = None
    for f in funcs:
        d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
        d2 = re.sub(r"\(int\)", "", d)
        d2 = re.sub(r"\s+", "", d2.split("(")[0].replace("void ", "").replace("b200sp::", ""))
        if d2 == want:
            target = f
            break
    if target is None:
        raise SystemExit("no cubin function for " + demangled)
    start = text.index("\n" + target + ":")
    end = text.index(".L_x_", text.index(".size", text.rindex(".global", 0, start)))  # unused; section end found below
    body = text[start:]
    nxt = body.find("//--------------------- .text.", 10)
    if nxt > 0:
        body = body[:nxt]
    cur = ("?", 0)
    out = []
    for ln in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s", ln):
            out.append((cur, ln.split("*/", 1)[1].strip()))
    return out


def main():
    rep, unit, needle = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    name, hdr, data = ncu_rows(rep, needle)
    ix = {h: i for i, h in enumerate(hdr)}
    sass = disasm_lines(unit, name)
    print("kernel:", name[:120])
    print("ncu SASS rows", len(data), "cubin instructions", len(sass))
    if len(data) != len(sass):
        print("WARNING: instruction counts differ: the object file is not the profiled build; lines are approximate")
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for i, r in enumerate(data):
        key = sass[i][0] if i < len(sass) else ("?", 0)
        a = agg[key]
        a[0] += int(r[ix["Instructions Executed"]])
        a[1] += int(r[ix["# Samples"]])
        a[2] += int(r[ix["L1 Wavefronts Shared"]] or 0)
        a[3] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    tw = sum(a[2] for a in agg.values())
    print(f"total warp-inst {ti}  samples {ts}  shared wavefronts {tw}")
    srcs = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        if f not in srcs:
            p = os.path.join(CSRC_DIR, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:70] if 0 < l <= len(srcs[f]) else ""
        print(f"{f}:{l:<5d} inst {100 * a[0] / ti:5.2f}%  samp {100 * a[1] / ts:5.2f}%  smem-wf {100 * a[2] / max(tw, 1):5.2f}% (excess {100 * a[3] / max(tw, 1):5.2f}%)  {text}")


if __name__ == "__main__":
    sys.exit(main())
