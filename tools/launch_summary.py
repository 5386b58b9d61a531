"""Turn an ncu launch-list CSV into the per-kernel summary + compact per-launch list kept under profiles/.

  gpurun -- 'ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none -c 4000 --csv \\
             --log-file gpurun_out/launches.csv python bench.py --quick --no-graph --steps 1 --warmup 3'
  python tools/launch_summary.py gpurun_out/launches.csv profiles/launches_rNN [--note "text for the header"]

The last complete training step is cut out between two `optim_kernel` launches.  Per-launch times under ncu are cold-cache
and serialised: compare shares, not absolutes."""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    lines = [l for l in open(src) if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        d = recs.setdefault(i, {"name": row["Kernel Name"], "grid": row["Grid Size"]})
        v = float(row["Metric Value"].replace(",", ""))
        unit, m = row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1000 if unit.startswith("n") else (v if unit.startswith("u") else v * 1000)
        else:
            d[m] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    ad = [i for i in recs if "optim_kernel" in recs[i]["name"] or "adamw" in recs[i]["name"]]
    if len(ad) < 2:
        raise SystemExit("need at least two optimizer launches to delimit a step (found %d)" % len(ad))
    lo, hi = ad[-2] + 1, ad[-1]
    step = [recs[i] for i in range(lo, hi + 1)]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in step:
        n = re.sub(r"\(.*", "", re.sub(r"<.*", "", d["name"])).replace("void ", "")
        a = agg[n]
        a[0] += 1
        a[1] += d["us"]
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    has_dram = any(a[2] for a in agg.values())
    out = ["# ncu launch list, one training step (launch IDs %d..%d = last complete step). %s" % (lo, hi, note),
           "# per-launch times are cold-cache and serialised: compare SHARES. total %.2f ms over %d launches" % (tot / 1e3, len(step)),
           "kernel,launches,ms,share" + (",dram_GB,dram_GB_per_s" if has_dram else "")]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = "%s,%d,%.3f,%.1f%%" % (n, a[0], a[1] / 1e3, 100 * a[1] / tot)
        if has_dram:
            line += ",%.3f,%.0f" % (a[2] / 1e9, a[2] / 1e9 / (a[1] / 1e6) if a[1] else 0)
        out.append(line)
    open(dst + "_summary.csv", "w").write("\n".join(out) + "\n")
    with open(dst + ".csv", "w") as f:
        f.write("id,kernel,grid,us,dram_read_MB,dram_write_MB\n")
        for i in range(lo, hi + 1):
            d = recs[i]
            f.write('%d,"%s","%s",%.2f,%.2f,%.2f\n' % (i, d["name"][:150], d["grid"], d["us"], d.get("dram__bytes_read.sum", 0) / 1e6,
                                                       d.get("dram__bytes_write.sum", 0) / 1e6))
    print("\n".join(out))
    gemm = agg.get("vlm::gemm_bf16_tcgen05_kernel")
    if gemm and has_dram:
        print("\nGEMM family DRAM traffic per step: %.3f GB over %d launches (profiles/gemm_traffic.json)" % (gemm[2] / 1e9, gemm[0]))


if __name__ == "__main__":
    main()
