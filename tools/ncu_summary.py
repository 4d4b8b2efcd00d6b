"""Turn ncu outputs into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches <launches.csv>
  python tools/ncu_summary.py full <report.ncu-rep>"""
import collections
import csv
import io
import subprocess
import sys

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("# per-kernel device time (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:")
    print("# compare SHARES, not absolutes).  %d launches, total %.1f us" % (sum(len(v) for v in agg.values()), tot / 1e3))
    print("%-100s %6s %12s %8s" % ("kernel", "n", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-100s %6d %12.2f %7.1f%%" % (k[:100], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
else:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    # pipe utilisation and the stall breakdown (which resource bounds the issue rate)
    extra = [c for c in h if ("pipe_" in c and "pct_of_peak" in c) or c.startswith("smsp__average_warps_issue_stalled")
             or c.startswith("smsp__average_warp_latency_issue_stalled") or c.startswith("smsp__warps_eligible")
             or c.startswith("smsp__issue_inst0") or c.startswith("smsp__inst_executed_pipe")
             or c.startswith("sm__inst_executed_pipe")]
    want = want + sorted(extra)
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")])
        for w in want:
            if w in h:
                print("  %-72s %18s %s" % (w, r[h.index(w)], units[h.index(w)]))
