"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/launch_summary.py launches.csv [out.md]"""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"(DeviceRadixSort\w+Kernel|DeviceScan\w*Kernel)", name)
    if m:
        return "cub::" + m.group(1)
    return name.split("(")[0].split("<")[0].strip()[-60:]


def main():
    lines = open(sys.argv[1]).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share | us / launch | grid | block |", "|---|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% | %.1f | %s | %s |" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0], v[2], v[3]))
    out.append("| total | %d | %.1f | | | | |" % (len(rows), tot))
    txt = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
