"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total us, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"]
        name = name.replace("mc::<unnamed>::", "mc::").replace("void ", "")
        name = name.split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print("%-72s %5s %10s %6s" % ("kernel", "count", "total us", "share"))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %5d %10.1f %5.1f%%" % (k, c, v, 100 * v / tot))
    print("%-72s %5d %10.1f" % ("TOTAL", sum(c for c, _ in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
