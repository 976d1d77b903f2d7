"""Developer tool: per-kernel counts of the Blackwell-native SASS instructions in the built library
(cuobjdump -sass): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor-map load), UBLKCP (bulk async
copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), UCGABAR (cluster barrier), plus registers where reported.

    python tools/sass_summary.py [path/to/libufv_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "STRONG.SYS", "FADD2", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ufvideo_b200", "libufv_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None:
            continue
        for k in KEYS:
            if re.search(r"\b" + k, line):
                counts[cur][k] += 1
        counts[cur]["instructions"] += 1 if re.match(r"\s+/\*[0-9a-f]{4}\*/", line) else 0
    names = demangle(order)
    print(f"# {os.path.relpath(lib, ROOT)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"# {'kernel':88s} " + " ".join(f"{k:>8s}" for k in ["instrs"] + KEYS))
    total = collections.Counter()
    for fn in order:
        c = counts[fn]
        short = re.sub(r"\(.*", "", names[fn]).replace("void ", "").replace("ufv::", "")
        print(f"{short[:90]:90s} " + " ".join(f"{c[k]:8d}" for k in ["instructions"] + KEYS))
        total.update(c)
    print(f"{'TOTAL':90s} " + " ".join(f"{total[k]:8d}" for k in ["instructions"] + KEYS))


if __name__ == "__main__":
    main()
