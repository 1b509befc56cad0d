"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def summarise(path, skip=()):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"^void ", "", name)
        name = re.split(r"[<(]", name)[0]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        if any(s in name for s in skip):
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"launches={sum(v[0] for v in agg.values())} total={tot / 1000:.3f} ms (cold-cache, serialised: compare shares)"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:45s} n={v[0]:6d} total={v[1] / 1000:9.3f} ms avg={v[1] / v[0]:9.2f} us share={100 * v[1] / tot:5.1f}%")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1], skip=tuple(sys.argv[2:])))
