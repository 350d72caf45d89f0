"""Summarise an ncu --metrics gpu__time_duration.sum launch list (CSV): per-kernel totals and one step's sequence."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def main():
    path = sys.argv[1]
    seq = len(sys.argv) > 2 and sys.argv[2] == "--seq"
    r = load(path)
    marks = [i for i, x in enumerate(r) if "build_keys" in x["Kernel Name"]]
    if len(marks) >= 3:
        r1 = r[marks[1]:marks[2]]
    else:
        r1 = r
    agg = collections.OrderedDict()
    tot = 0.0
    for x in r1:
        name = x["Kernel Name"].split("(")[0][:48]
        t = float(x["Metric Value"].replace(",", "")) / 1e3
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += t
        tot += t
    print("one step: %d launches, %.1f us serialized (cold-cache, per ncu)" % (len(r1), tot))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-50s n=%3d total=%8.1f us avg=%7.1f us share=%5.1f%%" % (k, n, t, t / n, 100 * t / tot))
    if seq:
        for x in r1:
            print("%-44s grid=%-16s block=%-13s %8.1f us" % (x["Kernel Name"].split("(")[0][:44], x["Grid Size"],
                                                             x["Block Size"], float(x["Metric Value"].replace(",", "")) / 1e3))


if __name__ == "__main__":
    main()
