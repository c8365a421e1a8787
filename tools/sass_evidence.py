"""Static evidence of what the kernels of librpo_b200.so are made of (no GPU needed): per kernel the ptxas
resource line (registers, spills, barriers) and the count of the SASS mnemonics that identify Blackwell-native code
(/opt/skills/guides/B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor
loads/stores, HMMA = legacy mma.sync, MUFU.EX2, FFMA2/FADD2 = packed f32x2.

    python tools/sass_evidence.py > profiles/r01_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpo_b200 import build  # noqa: E402

WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "LDGSTS", "MUFU.EX2",
         "FFMA2", "FADD2", "SYNCS", "UCGABAR", "ACQBULK"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(.*", "", name)            # drop the argument list
    name = name.replace("rpo::", "").replace("void ", "")
    return name


def main():
    objdir = os.path.join(os.path.dirname(build.LIB), "..", "build")
    rows = []
    for src in build.SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if not os.path.exists(obj):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for line in res.splitlines():
            m = re.match(r"\s*Function (\S+):", line)
            if m:
                cur = m.group(1)
            elif cur and "REG:" in line:
                usage[cur] = line.strip()
                cur = None
        counts = collections.defaultdict(collections.Counter)
        total = collections.Counter()
        fn = None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                fn = m.group(1)
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m and fn:
                total[fn] += 1
                op = m.group(1)
                for w in WATCH:
                    if op.startswith(w):
                        counts[fn][w] += 1
        names = demangle(list(total))
        for fn in total:
            rows.append((src, short(names.get(fn, fn)), total[fn], counts[fn], usage.get(fn, "")))
    print("# cuobjdump -sass / -res-usage of the sm_100a objects behind rpo_b200/lib/librpo_b200.so (tools/sass_evidence.py)")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, HMMA = mma.sync")
    for src, name, n, c, use in sorted(rows, key=lambda r: (r[0], r[1])):
        marks = " ".join(f"{k}={v}" for k, v in sorted(c.items()))
        use = re.sub(r"\s+", " ", use)
        print(f"{src:16s} {name[:86]:86s} sass={n:5d}  {marks}\n{'':16s}   {use}")


if __name__ == "__main__":
    main()
