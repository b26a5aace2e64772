"""Aggregate one C4 step from an ncu launch list (tools/bench_c4.py under ncu)."""
import collections, re, sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from launch_summary import load
seq = load(sys.argv[1])
idx = [i for i, (n, *_) in enumerate(seq) if "bert_embed_fwd" in n]
step = seq[idx[-1]:]
tot = sum(t for _, t, _, _ in step)
print(f"launches {len(step)}  sum of kernel times {tot/1e3:.2f} ms")
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t, g, b in step:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    agg[n][0] += 1; agg[n][1] += t
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print(f"{t/1e3:8.2f} ms {100*t/tot:5.1f}%  x{c:4d}  avg {t/c:8.1f} us  {n[:80]}")
