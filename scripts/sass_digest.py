"""Per-kernel SASS instruction digest of the built libsad_b200.so (runs anywhere: cuobjdump only).

    python scripts/sass_digest.py > profiles/sass_digest.txt

What to look for (B200_PROFILING.md): UTCHMMA / UTCHMMA.2CTA = tcgen05.mma (one / two SMs), LDTM = tcgen05.ld (TMEM -> registers),
UTMALDG = TMA tensor load, UBLKCP = TMA 1-D bulk copy, SYNCS = mbarrier, UTCBAR = tcgen05.commit, FFMA2 / FADD2 / FMUL2 = packed
fp32x2 arithmetic, MUFU = special-function unit, DFMA / DADD = fp64.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semi-supervised-adaptive-distillation_b200", "libsad_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "DFMA", "DADD",
        "F2F", "LDS", "STG", "LDG", "ATOM", "RED", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    kernels, cur = collections.OrderedDict(), None
    it = iter(names)
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(it)
            cur = re.sub(r"\([^()]*\)$", "", cur.strip())
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                    if k == "UTCHMMA" and ".2CTA" in op:
                        continue
                    kernels[cur][k] += 1
                    break
    print("# SASS digest of libsad_b200.so (scripts/sass_digest.py; static instruction counts per kernel, sm_100a)")
    print("# kernel | total | " + " ".join(KEYS))
    for name, c in kernels.items():
        cells = " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])
        print("%-110s total=%-6d %s" % (name[:110], c["total"], cells))


if __name__ == "__main__":
    sys.exit(main())
