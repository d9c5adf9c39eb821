"""Opcode digest of every kernel in libpm4b200.so and in the generated model modules (profiles/rNN_sass_digest.md):
   python tools/sass_digest.py > profiles/r02_sass_digest.md"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARK = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "HMMA", "FFMA2", "FADD2",
        "FMUL2", "FFMA", "DFMA", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOM", "RED", "BAR", "BRX"]


def digest(so):
    out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and name:
            kernels[name][m.group(1)] += 1
    return kernels


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    except OSError:
        return n


def main():
    sos = [os.path.join(ROOT, "pymc4_b200", "libpm4b200.so")] + sorted(glob.glob(os.path.join(ROOT, "pymc4_b200", "_kcache", "*.so")))
    want = sys.argv[1:]  # optional substrings of module file names
    print("# SASS digest (cuobjdump -sass, sm_100a): instruction count and marker opcodes per kernel\n")
    print("Markers: `UTCHMMA`/`LDTM`/`UTCBAR` = tcgen05 MMA / TMEM load / commit; `UBLKCP`+`SYNCS` = TMA bulk copy + mbarrier; "
          "`DMMA` = FP64 tensor-core MMA; `FFMA2`/`FADD2` = packed FP32 pairs (sm_100); `LDL`/`STL` = local memory (spills).\n")
    for so in sos:
        if want and not any(w in so for w in want):
            continue
        ks = digest(so)
        print("## %s\n" % os.path.relpath(so, ROOT))
        print("| kernel | instructions | markers |")
        print("|---|---|---|")
        for n, c in ks.items():
            total = sum(c.values())
            if total < 40:
                continue
            marks = ", ".join("%s x%d" % (k, c[k]) for k in MARK if c.get(k))
            d = demangle(n)
            d = re.sub(r"Model_[0-9a-f]{16}", "Model", d)
            print("| `%s` | %d | %s |" % (d[:150], total, marks))
        print()


if __name__ == "__main__":
    main()
