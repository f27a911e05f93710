"""Warp-stall samples of one profiled kernel per CUDA source line (read here, without a GPU).

  python tools/ncu_lines.py <report.ncu-rep> <demangled kernel substring> <mangled kernel substring> [launch index among the matches, default last] [min %]

ncu's CSV source page only carries SASS; the line of every instruction comes from `nvdisasm -g` on the cubin inside
argus_gui_b200/libargus_b200.so (built with -lineinfo), matched by instruction offset."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "argus_gui_b200", "libargus_b200.so")], cwd=tmp, capture_output=True)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, line, on = None, None, False
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                on = kernel in ln
                continue
            if not on:
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m:
                out[int(m.group(1), 16)] = line
        if out:
            break
    return out


def main():
    rep, kernel, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else -1
    minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kernel in r[1]]
    s = starts[which]
    e = min([i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > s] + [len(rows)])
    hdr, data = rows[s + 1], rows[s + 2:e]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    base = min(int(r[ix["Address"]], 16) for r in data)
    lines = sass_lines(mangled)
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = 0.0
    for r in data:
        n = float(r[ix["# Samples"]] or 0)
        tot += n
        ln = lines.get(int(r[ix["Address"]], 16) - base, ("?", 0))
        agg[ln]["n"] += n
        agg[ln]["inst"] += float(r[ix["Instructions Executed"]] or 0)
        for k in keys:
            agg[ln][k[6:]] += float(r[ix[k]] or 0)
    src = {}
    print("total samples %d" % tot)
    allst = collections.Counter()
    for ln, c in agg.items():
        for k, v in c.items():
            if k not in ("n", "inst"):
                allst[k] += v
    print("stall mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in allst.most_common(9)))
    for ln, c in sorted(agg.items(), key=lambda x: (x[0][0], x[0][1])):
        if c["n"] < minpct / 100 * tot:
            continue
        if ln[0] not in src:
            p = os.path.join(ROOT, "argus_gui_b200", "csrc", ln[0])
            src[ln[0]] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[ln[0]][ln[1] - 1].strip()[:80] if 0 < ln[1] <= len(src[ln[0]]) else ""
        st = ", ".join("%s %.0f%%" % (k, 100 * v / c["n"]) for k, v in c.most_common(5) if k not in ("n", "inst") and v > 0.12 * c["n"])
        print("%-22s %5.2f%%  %-80s | %s" % ("%s:%d" % ln, 100 * c["n"] / tot, text, st))


if __name__ == "__main__":
    main()
