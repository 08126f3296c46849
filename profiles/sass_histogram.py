#!/usr/bin/env python
"""SASS mnemonic histogram of every kernel in libboatenv.so (cuobjdump -sass), so that the evidence for the
TMA / mbarrier / tcgen05 claims is a committed file:  python profiles/sass_histogram.py > profiles/rNN_sass_histogram.txt

Per kernel: instruction count and the counts of the mnemonics that matter here
  UBLKCP      cp.async.bulk (TMA 1-D bulk copies, .S.G = global->shared, .G.S = shared->global)
  SYNCS       mbarrier operations (ARRIVE.TRANS64 / try_wait)
  UTCHMMA     tcgen05.mma          UTCBAR  tcgen05.commit      LDTM / STTM  tcgen05.ld / .st (TMEM)
  UTCATOMSWS  tcgen05.alloc / dealloc / relinquish
  LDG/STG/LDS/STS by width, F2I/I2F (fixed-point carriers), DFMA/DMUL/DADD (fp64 pipe), MUFU, VOTE, ATOM*/RED*
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sac-agent_b200", "libboatenv.so")
KEYS = ["UBLKCP", "SYNCS", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "LDG", "STG", "LDS", "STS", "LDGSTS",
        "F2I", "I2F", "F2F", "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "MUFU", "VOTE", "SHFL", "ATOM", "ATOMS", "RED",
        "BAR", "NANOSLEEP"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
            if op in ("LDG", "STG", "LDS", "STS", "UBLKCP"):
                kernels[cur][op + mods] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels (cuobjdump -sass, sm_100a)")
    for k, c in kernels.items():
        print(f"\n{names.get(k, k)}\n  instructions: {c['_total']}")
        row = [f"{key} {c[key]}" for key in KEYS if c[key]]
        print("  " + "  ".join(row))
        wide = [f"{key} {n}" for key, n in sorted(c.items()) if "." in key]
        if wide:
            print("  by form: " + "  ".join(wide))


if __name__ == "__main__":
    main()
