"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and print each kernel's share.

    python scripts/ncu_launch_shares.py profiles/r02v_ncu_launches_unit.csv.gz [first last]

`first last`: a launch-index window (default: everything).  ncu serialises kernels and flushes caches between them, so the
per-launch times are cold: the SHARE of a kernel must agree with the live run, not the absolute.
"""
import csv
import gzip
import io
import re
import sys
from collections import defaultdict

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rows = list(csv.reader(io.StringIO("".join(lines))))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
recs = [(r[ki], float(r[vi].replace(",", "")) / 1e6) for r in rows[1:] if len(r) > vi]
first, last = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(recs))
recs = recs[first:last]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("dv::<unnamed>::", "").replace("void ", "")
    return name if len(name) < 48 else name[:45] + "..."


agg = defaultdict(lambda: [0, 0.0])
for k, ms in recs:
    a = agg[short(k)]
    a[0] += 1
    a[1] += ms
total = sum(a[1] for a in agg.values())
print(f"# {path}: launches [{first}, {last}) = {len(recs)} launches, {total:.2f} ms under ncu")
print(f"{'kernel':50s} {'n':>6s} {'ms':>9s} {'share':>7s} {'us/launch':>10s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if ms / total < 0.002:
        continue
    print(f"{k:50s} {n:6d} {ms:9.3f} {100 * ms / total:6.1f}% {1e3 * ms / n:10.2f}")
