#!/usr/bin/env python
"""Static evidence from the built objects (no GPU needed): per kernel, the ptxas resource line (registers, spills, shared memory)
and a histogram of the SASS mnemonics that carry the design claims -- packed FP32 (FFMA2 / FADD2 / FMUL2), MUFU.RSQ / MUFU.RCP,
128-bit global and shared accesses, warp votes / shuffles / MATCH, atomics.

    python profiles/static_evidence.py > profiles/r02_static_sass_ptxas.txt
Needs planetmodel-sph_b200/csrc/*.o (make -C planetmodel-sph_b200/csrc) and the CUDA toolkit's cuobjdump / cu++filt."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "planetmodel-sph_b200", "csrc")
CUDA = "/usr/local/cuda/bin"
OBJS = ["kernels_gravity", "kernels_tree", "kernels_neighbors", "kernels_sph", "kernels_sort", "kernels_grid", "kernels_group"]
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FMNMX", "MUFU.RSQ", "MUFU.RCP", "MUFU.SQRT", "LDG.E.128", "LDG.E.64", "LDG.E",
         "STG.E.128", "STG.E", "LDS.128", "LDS.64", "LDS", "STS", "SHFL", "VOTE", "MATCH", "REDUX", "ATOMS", "ATOMG", "RED", "BAR", "BSSY"]


def demangle(name):
    try:
        out = subprocess.check_output([os.path.join(CUDA, "cu++filt"), name]).decode().strip()
    except Exception:
        out = name
    out = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", out)
    out = re.sub(r"\((?:int|bool)\)", "", out)            # template arguments print as (int)6, (bool)1
    return out.split("(")[0].replace("void ", "")


def main():
    for obj in OBJS:
        path = os.path.join(CSRC, obj + ".o")
        res = subprocess.check_output([os.path.join(CUDA, "cuobjdump"), "-res-usage", path], stderr=subprocess.STDOUT).decode()
        usage = {}
        for m in re.finditer(r"Function (\S+):\n\s*(REG:\d+[^\n]*)", res):
            usage[m.group(1)] = m.group(2).strip()
        sass = subprocess.check_output([os.path.join(CUDA, "cuobjdump"), "-sass", path]).decode()
        print("== %s.cu" % obj)
        for blk in re.split(r"\n\s*Function : ", sass)[1:]:
            name = blk.split("\n", 1)[0].strip()
            ops = collections.Counter()
            total = 0
            for line in blk.splitlines():
                m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                if not m:
                    continue
                total += 1
                op = m.group(1)
                for w in WATCH:
                    if op == w or op.startswith(w + "."):
                        ops[w] += 1
                        break
            hist = "  ".join("%s %d" % (w, ops[w]) for w in WATCH if ops[w])
            print("  %-44s %5d SASS instructions | %s" % (demangle(name)[:44], total, usage.get(name, "")))
            print("      " + hist)


if __name__ == "__main__":
    main()
