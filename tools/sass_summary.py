#!/usr/bin/env python3
"""Opcode histogram per kernel of the built librexsim.so (cuobjdump -sass), written to profiles/<tag>_sass_summary.txt, so the
Blackwell-native claims (TMA bulk copies: UBLKCP + SYNCS mbarrier waits; packed fp32x2 FFMA2 in the policy kernel; local-memory
traffic LDL/STL; fp64 DFMA/DMUL of the gait timing; tensor-core UTC*MMA / LDTM if any) do not depend on someone disassembling
the library.  Usage: python tools/sass_summary.py r02   (runs here: no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "rex_gym_b200", "librexsim.so")
WATCH = ["UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "FFMA2", "FFMA", "FMUL", "FADD",
         "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "LDL", "STL", "LDS", "STS", "LDG", "STG", "BAR", "BSSY", "BRA"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    with open(path, "w") as f:
        f.write("# cuobjdump -sass rex_gym_b200/librexsim.so (sm_100a): static instruction counts per kernel, by opcode stem.\n")
        f.write("# columns: total | " + " ".join(WATCH) + "\n")
        for (k, c), name in zip(kernels.items(), demangle):
            name = re.sub(r"\(rexsim::Params\)|\(.*\)$", "", name)
            tot = sum(c.values())
            f.write(f"{name}\n    total {tot}  " + "  ".join(f"{w} {c[w]}" for w in WATCH if c[w]) + "\n")
    print("wrote", path, len(kernels), "kernels")


if __name__ == "__main__":
    main()
