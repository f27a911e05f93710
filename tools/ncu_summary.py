"""Summarise ncu outputs (read here, without a GPU) into small text files for profiles/.

  python tools/ncu_summary.py launches <launches.csv>            per-kernel totals and shares of a gpu__time_duration launch list
  python tools/ncu_summary.py full <report.ncu-rep>               key metrics + stall reasons of every profiled launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] in ("us", "usecond") else v)
        name = r[ki].split("(")[0].replace("void ", "")
        agg[name][0] += v; agg[name][1] += 1
    tot = sum(a[0] for a in agg.values())
    print("kernel                                   launches   total ms   ms/launch   share")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-40s %8d %10.3f %11.4f %6.1f%%" % (k[:40], a[1], a[0], a[0] / a[1], 100 * a[0] / tot))
    print("total %.3f ms over %d launches (cold-cache, serialised: compare shares, not absolutes)" % (tot, sum(a[1] for a in agg.values())))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")][:140])
        for k in KEYS:
            if k in hdr:
                print("  %-80s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(h, float(r[hdr.index(h)].replace(",", "") or 0)) for h in hdr
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        print("  stall reasons (warps per issue): " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                                                                for h, v in sorted(st, key=lambda x: -x[1])[:7]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
