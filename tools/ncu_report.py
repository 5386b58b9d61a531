"""Summarise one `ncu --set full --import-source on` report (first kernel in it): headline metrics, instruction mix by opcode
and the stall attribution of the most-sampled SASS instructions — the analysis behind profiles/ncu_*.txt.

  gpurun -- 'ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/x -f python tools/...'
  python tools/ncu_report.py gpurun_out/x.ncu-rep [--elements N]      (N: output / score elements, for instructions per element)
"""
import collections
import csv
import io
import subprocess
import sys

HEADLINE = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    elements = float(sys.argv[sys.argv.index("--elements") + 1]) if "--elements" in sys.argv else None
    raw = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    kv = dict(zip(hdr, zip(units, vals)))
    print("kernel:", kv.get("Kernel Name", ("", "?"))[1])
    for k in HEADLINE:
        if k in kv:
            print("  %-66s %-16s %s" % (k, kv[k][0], kv[k][1]))
    rows = list(csv.reader(io.StringIO(page(rep, "source"))))
    h, data = rows[1], rows[2:]
    ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    byop, tot, st = collections.Counter(), 0, collections.Counter()
    for r in data:
        if len(r) <= isamp:
            continue
        t = r[ia].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        n = int(r[ie] or 0)
        byop[op] += n
        tot += n
        for i in stalls:
            st[h[i]] += int(r[i] or 0)
    print("warp instructions: %.2f M%s" % (tot / 1e6, "  (%.1f per element)" % (tot * 32 / elements) if elements else ""))
    print("  mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in byop.most_common(12)))
    print("  stall samples:", ", ".join("%s %d" % (k[6:], v) for k, v in st.most_common(7)))
    print("  most-sampled instructions:")
    for r in sorted([r for r in data if len(r) > isamp and r[isamp]], key=lambda r: -int(r[isamp]))[:14]:
        top = sorted(((int(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
        print("    %s %5s  %-64s %s" % (r[0][-5:], r[isamp], r[ia][:64], top))


if __name__ == "__main__":
    main()
