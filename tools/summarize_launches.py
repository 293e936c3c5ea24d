"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, average, share)."""
import collections
import csv
import io
import sys


def main(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.DictReader(io.StringIO("".join(rows))))
    agg = collections.OrderedDict()
    for x in r:
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e6 if u == "ns" else v / 1e3 if u.startswith("us") else v
        a = agg.setdefault(x["Kernel Name"].split("(")[0][:48], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(r)} launches, {tot:.3f} ms total (ncu per-launch times are cold-cache and serialised: compare shares)")
    print(f"{'kernel':48s} {'launches':>8s} {'total ms':>10s} {'avg ms':>9s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} {n:8d} {t:10.3f} {t / n:9.4f} {t / tot * 100:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
