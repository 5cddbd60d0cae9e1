"""SASS opcode histogram per kernel of libmrfp_b200.so (cuobjdump -sass), written to profiles/sass_opcodes.txt.

    python tools/sass_hist.py [--all]

Lists, for every kernel, the tensor-core / TMA / tensor-memory / matrix-move opcodes that prove which hardware path it
uses (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, LDTM / STTM =
tcgen05.ld / st, LDSM / STSM = ldmatrix / stmatrix, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and the total
instruction count; --all adds the full histogram.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mrfp_b200", "lib", "libmrfp_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDTM", "STTM", "LDSM", "STSM", "SYNCS", "UTMAPF",
       "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDGSTS", "HMMA", "DADD", "DFMA", "MUFU")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    dm = demangle(list(kernels))
    lines = ["# SASS opcode histogram of mrfp_b200/lib/libmrfp_b200.so (sm_100a), `python tools/sass_hist.py`",
             "# kernel | total instructions | key opcodes", ""]
    for k, c in kernels.items():
        name = re.sub(r"\(anonymous namespace\)::", "", dm.get(k, k))
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        keys = ", ".join(f"{o} {c[o]}" for o in KEY if c.get(o))
        lines.append(f"{name} | {sum(c.values())} | {keys}")
        if "--all" in sys.argv:
            lines.append("    " + ", ".join(f"{o} {n}" for o, n in c.most_common()))
    out = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
