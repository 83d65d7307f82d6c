#!/usr/bin/env python
"""Per-kernel SASS opcode counts of the in-tree library (what proves the Blackwell-native path, B200_PROFILING.md):
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync (must be absent).   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "va_vae_b200", "libvavae_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2", "LDG", "STG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for key in OPS:
                if op == key or op.startswith(key + ".") or (key.count(".") and key in op):
                    cur[key] += 1
    names = demangle(list(per))
    total = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); opcode counts per kernel")
    print("# " + " ".join(f"{k:>13}" for k in ["instrs"] + OPS) + "  kernel")
    for fn, c in per.items():
        total.update(c)
        full = names.get(fn, fn)
        cut = full.find(">(")            # templates print "(int)5" inside <...>: cut at the parameter list, not there
        short = full[: cut + 1] if cut >= 0 else re.sub(r"\(.*", "", full)
        short = short.replace("void ", "").replace("(int)", "").replace("(bool)", "")
        print("  " + " ".join(f"{c[k]:>13}" for k in ["_total"] + OPS) + "  " + short[:110])
    print("# " + " ".join(f"{total[k]:>13}" for k in ["_total"] + OPS) + "  TOTAL")
    if total["HMMA"]:
        sys.exit("legacy HMMA found")


if __name__ == "__main__":
    main()
