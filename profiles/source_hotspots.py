"""Rank CUDA source lines of an ncu report (collected with --import-source on, code built with -lineinfo) by executed
warp instructions and stall samples:  python profiles/source_hotspots.py report.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    fname, cur = None, None
    per = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif r[0].isdigit() and len(r) > ii:
            try:
                n, s = int(r[ii]), int(r[isamp])
            except ValueError:
                continue
            key = (fname, int(r[0]))
            if key not in per:
                per[key] = [0, 0, r[1].strip()[:120]]
            per[key][0] += n
            per[key][1] += s
    tot = sum(v[0] for v in per.values()) or 1
    tots = sum(v[1] for v in per.values()) or 1
    print("total warp instructions %d, samples %d" % (tot, tots))
    for (f, l), (n, s, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%% inst %6.2f%% samp  %s:%d  %s" % (100.0 * n / tot, 100.0 * s / tots, f, l, src))


if __name__ == "__main__":
    main(sys.argv[1])
