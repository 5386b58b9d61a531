"""Per-CUDA-source-line totals of one ncu report (needs -lineinfo + --import-source on): executed warp instructions and stall samples.
  python tools/ncu_lines.py gpurun_out/x.ncu-rep [--top N] [--file attention_tc.cu]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    want = sys.argv[sys.argv.index("--file") + 1] if "--file" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue                                  # SASS rows have an empty line number: the CUDA row already carries their sum
        if want and (cur_file is None or want not in cur_file):
            continue
        # source text with embedded quotes (inline asm) splits into extra fields: index the numeric columns from the END
        n = len(hdr)
        ie, isamp = hdr.index("Instructions Executed") - n, hdr.index("# Samples") - n
        stalls = [(i - n, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        key = (cur_file.split("/")[-1], int(r[0]))
        ent = lines.setdefault(key, [0, 0, {}, r[1]])
        ent[0] += int(r[ie] or 0)
        ent[1] += int(r[isamp] or 0)
        for i, nm in stalls:
            v = int(r[i] or 0)
            if v:
                ent[2][nm] = ent[2].get(nm, 0) + v
    tot_i = sum(e[0] for e in lines.values())
    tot_s = sum(e[1] for e in lines.values())
    print("total warp instructions %.2f M, samples %d" % (tot_i / 1e6, tot_s))
    print("--- by instructions executed")
    for k, e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-18s %5d  %6.2f%% inst %6.2f%% smp  %s" % (k[0][:18], k[1], 100.0 * e[0] / max(tot_i, 1), 100.0 * e[1] / max(tot_s, 1), e[3].strip()[:110]))
    print("--- by stall samples")
    for k, e in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(e[2].items(), key=lambda kv: -kv[1])[:3]
        print("%-18s %5d  %6.2f%% smp %6.2f%% inst  %-70s %s" % (k[0][:18], k[1], 100.0 * e[1] / max(tot_s, 1), 100.0 * e[0] / max(tot_i, 1), e[3].strip()[:70], st))


if __name__ == "__main__":
    main()
