"""Summarise one training step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    seq = []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        t = t / 1e3 if u == "ns" else t * 1e3 if u == "ms" else t
        seq.append((row["Kernel Name"], t, row.get("Grid Size"), row.get("Block Size")))
    return seq


def main(path, top=28):
    seq = load(path)
    idx = [i for i, (n, *_r) in enumerate(seq) if "embedding_fwd" in n]
    # the LAST complete step of the list (the first one also holds the workspace allocation fills)
    a, b = (idx[-2], idx[-1]) if len(idx) > 1 else (idx[0], len(seq))
    step = seq[a:b]
    tot = sum(t for _, t, _, _ in step)
    print(f"launches in step: {len(step)}   sum of kernel times: {tot:.1f} us")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, g, bk in step:
        key = re.sub(r"\(.*", "", n)[:64]
        agg[key][0] += 1
        agg[key][1] += t
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{c:3d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
