"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/summarize_launches.py profiles/r1a/ncu_launches_config2_2steps.csv
"""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
        k = d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms summed (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e6:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg {v[1] / v[0] / 1e3:8.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
