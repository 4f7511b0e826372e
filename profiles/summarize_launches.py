"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, total, share."""
import collections
import csv
import sys


def summarize(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = row["Kernel Name"][:90]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["%d launches, %.1f us total (cold-cache, serialised: compare shares, not absolutes)" % (sum(v[0] for v in agg.values()), tot)]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append("%10.1f us avg  x%4d  %6.2f%%  %s" % (t / c, c, 100 * t / tot, k))
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
