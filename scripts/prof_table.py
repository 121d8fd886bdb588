"""Group the per-launch CSV written by `bench.py --profile-dump` by launch tag.

    python scripts/prof_table.py gpurun_out/prof.csv [top_n]
"""
import collections
import csv
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault((r["kind"], r["tag"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += float(r["ms"])
    a[2] += float(r["flops"])
    a[3] += float(r["bytes"])
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms")
bykind = collections.defaultdict(float)
for (k, _), a in agg.items():
    bykind[k] += a[1]
print("by kind (0 gemm, 1 conv, 2 attn, 3 other):", {k: round(v, 2) for k, v in sorted(bykind.items())})
print(f"{'tag':58s} {'n':>5s} {'ms':>8s} {'us/launch':>9s} {'TFLOP/s':>8s} {'GB/s':>7s}")
for (k, tag), (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    tf = fl / 1e12 / (ms / 1e3) if ms > 0 else 0
    gb = by / 1e9 / (ms / 1e3) if ms > 0 else 0
    print(f"{tag:58s} {n:5d} {ms:8.3f} {ms * 1e3 / n:9.1f} {tf:8.1f} {gb:7.0f}")
