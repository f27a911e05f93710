"""Per-kernel SASS instruction counts of libargus_b200.so (cuobjdump, no GPU needed): the mnemonics that show which hardware
paths a kernel uses - DMMA (fp64 tensor pipe), UBLKCP (bulk-async / TMA copies), SYNCS (mbarrier), UCGABAR (cluster barrier),
LDGSTS (cp.async), DFMA / DMUL / DADD, shuffles, spills.    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "argus_gui_b200", "libargus_b200.so")
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "UBLKCP", "SYNCS", "UCGABAR", "LDGSTS", "SHFL", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    counts, total, name = collections.OrderedDict(), collections.Counter(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("argus::", "")
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for k in KEYS:
                if op == k or op.startswith(k):
                    counts[name][k] += 1
                    break
    print("libargus_b200.so: cubins for %s only; %d kernels, %d instructions" % (", ".join(arch), len(counts), sum(total.values())))
    print("%-34s %7s " % ("kernel", "instr") + " ".join("%7s" % k for k in KEYS))
    for n, c in sorted(counts.items(), key=lambda kv: -total[kv[0]]):
        print("%-34s %7d " % (n[:34], total[n]) + " ".join("%7s" % (c[k] if c[k] else ".") for k in KEYS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("%-34s %7d " % ("all kernels", sum(total.values())) + " ".join("%7s" % (tot[k] if tot[k] else ".") for k in KEYS))
    print("no UTCMMA / LDTM / UTMALDG: tcgen05 has no fp64 kind - the fp64 tensor path is DMMA (mma.sync.m8n8k4.f64)")


if __name__ == "__main__":
    main()
