"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total time, share of the step)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    txt = [f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms total (ncu per-launch times: cold cache, serialised)",
           f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        txt.append(f"{k:62s} {v[0]:8d} {v[1] / 1e6:10.3f} {100 * v[1] / tot:6.1f}% {v[1] / v[0] / 1e3:9.1f}")
    s = "\n".join(txt) + "\n"
    if out:
        open(out, "w").write(s)
    else:
        sys.stdout.write(s)


if __name__ == "__main__":
    main(*sys.argv[1:3])
