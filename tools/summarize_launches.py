"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [n_forwards] > profiles/xxx.md"""
import csv, re, sys
from collections import defaultdict

path = sys.argv[1]
nfwd = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name)
    agg[name][0] += 1
    agg[name][1] += float(r[iv]) / 1e6
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith("gf::") or k.startswith("fl::"))
print(f"# ncu launch list summary: {path}\n")
print(f"{len(rows) - 1} launches over {nfwd:g} forward passes of one batch (16 pairs, 640x480, dense regime); "
      f"per-launch times are cold-cache and serialised (compare SHARES, not absolutes).\n")
print(f"total {tot / nfwd:.2f} ms per forward; kernels of libgeoformer_sm100.so (`gf::`, `fl::` = `gf::fl::`): {100 * ours / tot:.1f} % of GPU time\n")
print("| kernel | launches/fwd | ms/fwd | share |\n|---|---|---|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| `{k[:110]}` | {c / nfwd:.1f} | {ms / nfwd:.3f} | {100 * ms / tot:.1f} % |")
