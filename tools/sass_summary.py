"""Count the SASS mnemonics that prove the tcgen05 / TMEM / TMA path per kernel of libvlmb200.so (B200_PROFILING.md "SASS mnemonics").
Usage: python tools/sass_summary.py > profiles/sass_summary.txt       (CPU only: cuobjdump on the built library)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vilmedic_b200", "libvlmb200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "HMMA", "MUFU.EX2", "STL", "LDL",
         "RED.E.ADD", "ATOMG", "LDG.E.128", "STG.E.128"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    cur[w] += 1
    print("# SASS mnemonic counts per kernel of vilmedic_b200/libvlmb200.so (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma,")
    print("# LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
    print("# HMMA = legacy mma.sync, STL/LDL = local-memory (spill) traffic.  Template instances of one kernel are merged (max per column).")
    merged = collections.OrderedDict()
    for name, c in per.items():
        base = re.sub(r"<.*", "", name).replace("void ", "").replace("vlm::", "")
        m = merged.setdefault(base, [0, collections.Counter()])
        m[0] += 1
        for k, v in c.items():
            m[1][k] = max(m[1][k], v)
    cols = [w for w in WATCH if any(m[1][w] for m in merged.values())]
    print("%-34s %5s %7s " % ("kernel", "inst.", "SASS") + " ".join("%8s" % c[:8] for c in cols))
    for base, (n, c) in sorted(merged.items(), key=lambda kv: -kv[1][1]["UTCHMMA"] * 100000 - kv[1][1]["_total"]):
        print("%-34s %5d %7d " % (base[:34], n, c["_total"]) + " ".join("%8d" % c[w] for w in cols))


if __name__ == "__main__":
    sys.exit(main())
