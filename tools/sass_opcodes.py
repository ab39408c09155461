#!/usr/bin/env python
"""Per-kernel SASS opcode summary of rmem_b200/lib/librmem_b200.so (cuobjdump -sass): which kernels are Blackwell-native
(tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG, tcgen05.commit = UTCBAR) and which still use the legacy
tensor path (mma.sync = HMMA).  Writes profiles/r02_sass_opcodes.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rmem_b200", "lib", "librmem_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTCATOMSWS", "HMMA",
       "MUFU.EX2", "USETMAXREG", "UCGABAR_ARV", "SYNCS", "LDGSTS", "BAR.SYNC", "BAR.ARV"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    kernels[cur][o] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                kernels[cur]["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and ".2CTA" in op:
                kernels[cur]["UTMALDG.2CTA"] += 1
            if op.startswith("UTCBAR") and ".2CTA" in op:
                kernels[cur]["UTCBAR.2CTA"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "HMMA", "MUFU.EX2", "USETMAXREG", "LDGSTS"]
    lines = [f"SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)",
             "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, "
             "UTMALDG = cp.async.bulk.tensor (TMA), HMMA = mma.sync (legacy tensor path), USETMAXREG = setmaxnreg", "",
             f"{'kernel':58s} " + " ".join(f"{c:>12s}" for c in cols) + f" {'instrs':>8s}"]
    for (k, cnt), name in zip(kernels.items(), demangle):
        short = name.replace("(anonymous namespace)::", "").replace("rmem::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        lines.append(f"{short[:58]:58s} " + " ".join(f"{cnt.get(c, 0):12d}" for c in cols) + f" {cnt['_total']:8d}")
    txt = "\n".join(lines) + "\n"
    dst = os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
    open(dst, "w").write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
