"""Aggregate an ncu source page (cuda,sass view) per CUDA source line: stall samples and warp instructions executed.
usage: python tools/ncu_hotspots.py report.ncu-rep 'regex:k_sdf_fused_backward' [launch-skip] [top]
(reads a .ncu-rep brought back from the GPU box; runs on the CPU container)"""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                          "::%s:%s" % (kern, str(int(skip) + 1))], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, agg, hdr = None, {}, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0] not in ("", "Function Name") and hdr and len(r) >= 8 and r[0].isdigit():
            si, ni, ii = hdr.index("# Samples"), hdr.index("Warp Stall Sampling (Not-issued Samples)"), hdr.index("Instructions Executed")
            key = (cur_file, int(r[0]))
            s = agg.setdefault(key, [0, 0, 0, r[1]])
            s[0] += num(r[si]); s[1] += num(r[ni]); s[2] += num(r[ii])
    tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[2] for v in agg.values())
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    print("%7s %7s %10s  %s" % ("samples", "notiss", "warp-inst", "file:line source"))
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%7d %7d %10d  %s:%d %s" % (v[0], v[1], v[2], f, ln, v[3].strip()[:110]))
    print("--- by instructions")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%7d %7d %10d  %s:%d %s" % (v[0], v[1], v[2], f, ln, v[3].strip()[:110]))


if __name__ == "__main__":
    main()
