"""Per-kernel SASS opcode evidence: counts of the mnemonics that prove (or disprove) a Blackwell-native kernel
(B200_PROFILING.md "What proves a Blackwell-native kernel") for every kernel in libscot_b200.so.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store /
reduce, UBLKCP = bulk copy, HMMA = legacy mma.sync, SYNCS = mbarrier, MUFU = special-function unit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "poseidon_b200", "libscot_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "SYNCS", "MUFU", "LDGSTS",
        "ATOMS", "RED", "ATOMG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    cur[k] += 1
    dm = demangle(list(kernels))
    print(f"# SASS opcode summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), {len(kernels)} kernels")
    print("# " + " ".join(f"{k:>8s}" for k in ["instrs"] + KEYS) + "  kernel")
    rows = []
    for name, c in kernels.items():
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", dm.get(name, name))
        short = re.sub(r"\(.*\)$", "", short)
        rows.append((short, c))
    for short, c in sorted(rows):
        print("  " + " ".join(f"{c[k]:8d}" for k in ["_total"] + KEYS) + "  " + short[:110])
    tc = sorted(s for s, c in rows if c["UTCHMMA"] or c["UTCQMMA"])
    legacy = sorted(s for s, c in rows if c["HMMA"] and not c["UTCHMMA"])
    print(f"\n# {len(tc)} kernels issue tcgen05.mma (UTC*MMA); {len(legacy)} kernels use legacy mma.sync (HMMA) only:")
    for s in legacy:
        print("#   legacy:", s[:110])


if __name__ == "__main__":
    main()
