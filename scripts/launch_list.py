"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) per kernel: launches, time, share.
usage: python scripts/launch_list.py launches.csv "command that was profiled" """
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
    t = collections.defaultdict(float)
    c = collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4])
        t[name] += float(r[14].replace(",", ""))
        c[name] += 1
    tot = sum(t.values())
    print("ncu launch list of: %s" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
    print("(ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)")
    print("unit: ns   total %.3f   launches %d\n" % (tot, len(rows)))
    for k in sorted(t, key=t.get, reverse=True):
        print("%-90s launches %4d  time %14.1f  share %5.1f%%" % (k, c[k], t[k], 100 * t[k] / tot))


if __name__ == "__main__":
    main()
