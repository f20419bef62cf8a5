"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of device time)."""
import collections
import csv
import sys


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[r[ki][:100]][0] += 1
        agg[r[ki][:100]][1] += v
        tot += v
    print(f"total {tot:.1f} us over {len(data)} launches")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}% n={c:4d} avg={t / c:7.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
