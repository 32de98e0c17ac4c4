"""Per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / bulk-copy use (the mnemonics of /opt/skills/guides/
B200_PROFILING.md) in the built library:  python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "alphazero_risk_b200", "libaz_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "IMMA",
       "FFMA", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "SHFL", "VOTE", "POPC", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    name = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            per[name]["_total"] += 1
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + ".") or op.startswith(o + "_"):
                    per[name][o] += 1
                    break
    sha = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    print("# cuobjdump -sass alphazero_risk_b200/libaz_b200.so (sha256 %s...), sm_100a; instruction counts per kernel" % sha)
    print("# UTCHMMA = tcgen05.mma (bf16), LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,")
    print("# UTMALDG = tensor-map TMA load (not used: operands are contiguous slabs), SYNCS = mbarrier ops")
    cols = [o for o in OPS if any(c[o] for c in per.values())]
    print("%-44s %7s " % ("kernel", "instrs") + " ".join("%7s" % o for o in cols))
    for k, c in per.items():
        print("%-44s %7d " % (k[:44], c["_total"]) + " ".join("%7d" % c[o] for o in cols))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("%-44s %7d " % ("TOTAL (%d kernels)" % len(per), tot["_total"]) + " ".join("%7d" % tot[o] for o in cols))


if __name__ == "__main__":
    sys.exit(main())
