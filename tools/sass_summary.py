#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library: counts of the Blackwell-specific instructions
(UTCHMMA / UTCHMMA.2CTA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA / legacy mma.sync), registers and spills.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (no GPU needed: cuobjdump reads the .so)
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "video-moment-localization_b200", "libvml_b200.so")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
PATTERNS = OrderedDict([("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"),
                        ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                        ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("SHFL", r"\bSHFL"), ("FFMA2/FADD2/FMUL2", r"\bF(FMA|ADD|MUL)2\b"),
                        ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("LDS", r"\bLDS"), ("STS", r"\bSTS")])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    kernels, cur = OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        kernels[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
        for key, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][key] += 1
    names = demangle(list(kernels))
    total = Counter()
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass / -res-usage, sm_100a)")
    print(f"# {len(kernels)} kernels; columns: " + ", ".join(PATTERNS))
    for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["UTCHMMA"] - kv[1]["UTCHMMA.2CTA"] - kv[1]["UTMALDG"]):
        short = re.sub(r"\(.*", "", names.get(k, k)).replace("vml::", "").replace("void ", "")
        reg, shm, loc = usage.get(k, (0, 0, 0))
        cols = " ".join(f"{key}={c[key]}" for key in PATTERNS if c[key])
        print(f"{short[:90]:90s} insts={c['instructions']:6d} regs={reg:3d} local={loc:4d}  {cols}")
        total.update(c)
    print("# totals: " + " ".join(f"{key}={total[key]}" for key in PATTERNS))


if __name__ == "__main__":
    sys.exit(main())
