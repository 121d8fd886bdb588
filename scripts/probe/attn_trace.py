"""Probe (GPU): clock64 timeline of the heaviest CTA of the two-query-tile attention kernel (attn_pair_kernel).

Builds a second copy of the library with -DDV_ATTN_TRACE (attention.cu only; the other objects are the product's), runs
one attention launch at a benchmark shape and prints, per key tile j, where a softmax warp of each group and the MMA
thread spent their cycles.
    python scripts/probe/attn_trace.py [B L H]
"""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
here = Path(__file__).resolve().parent
poly = next((a.split("=")[1] for a in sys.argv if a.startswith("--poly=")), None)
so = here / (f"libdeepv_trace_p{poly}.so" if poly is not None else "libdeepv_trace.so")
csrc = ROOT / "deepv_b200" / "csrc"
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]
if not so.exists() or "--rebuild" in sys.argv:
    from deepv_b200 import build
    build.build(force=not list((ROOT / "deepv_b200" / "build").glob("*.o")))   # (the object files do not travel with gpurun)
    obj = here / f"attention_trace_{poly}.o"
    extra = [f"-DDV_ATTN_POLY_PAIRS={poly}"] if poly is not None else []
    subprocess.check_call(["nvcc", *flags, "-DDV_ATTN_TRACE", *extra, "-c", str(csrc / "attention.cu"), "-o", str(obj)])
    others = [str(o) for o in (ROOT / "deepv_b200" / "build").glob("*.o") if o.name != "attention.o"]
    subprocess.check_call(["nvcc", "-shared", "-o", str(so), str(obj), *others, "-gencode", "arch=compute_100a,code=sm_100a",
                           "-cudart", "static"])
if "--build-only" in sys.argv:
    sys.exit(0)

os.environ.setdefault("DV_ATTN_PIPE", "2")
lib = C.CDLL(str(so))
lib.dv_attn_trace_buffer.restype = C.c_void_p
lib.dv_attention.restype = C.c_int
lib.dv_attention.argtypes = [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p]
args = [int(x) for x in sys.argv[1:] if x.isdigit()]
B, L, H = (args + [3, 2237, 24])[:3] if len(args) >= 3 else (3, 2237, 24)
torch.manual_seed(0)
qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
kv = torch.full((L,), L, dtype=torch.int32, device="cuda")       # every key visible: the last CTA walks all tiles
Lpad = (L + 127) // 128 * 128
kb = torch.zeros(B, Lpad, device="cuda")
kb[:, L:] = float("-inf")
out = torch.empty(B, L, H * 64, device="cuda", dtype=torch.bfloat16)
buf = lib.dv_attn_trace_buffer()
for _ in range(3):
    rc = lib.dv_attention(qkv.data_ptr(), out.data_ptr(), kv.data_ptr(), kb.data_ptr(), B, L, Lpad, H, None)
    assert rc == 0, rc
torch.cuda.synchronize()
import numpy as np  # noqa: E402
n_cta = H * B * ((L + 127) // 128)
tr = np.zeros(1024 + 8 * 8192, dtype=np.int64)
cudart = C.CDLL("libcudart.so.12")
assert cudart.cudaMemcpy(tr.ctypes.data_as(C.c_void_p), C.c_void_p(buf), C.c_size_t(tr.nbytes), 2) == 0
per_cta = tr[1024:1024 + 8 * min(n_cta, 8192)].reshape(-1, 8)
tr = tr[:1024].reshape(4, 32, 8)
if per_cta[:, 1].any():
    # per-CTA phases (ns, globaltimer) and the idle gaps between consecutive CTAs of an SM slot
    ph = per_cta[per_cta[:, 1] > 0]
    names_c = ["entry->setup", "setup->Q parked", "Q parked->first S", "main loop", "last P->stored", "stored->exit"]
    d = np.diff(ph[:, 1:8], axis=1)
    print("== per-CTA phases, ns (mean / median / p90) over", len(ph), "CTAs")
    for i, nm in enumerate(names_c):
        print(f"  {nm:22s} {d[:, i].mean():9.0f} {np.median(d[:, i]):9.0f} {np.percentile(d[:, i], 90):9.0f}")
    print(f"  whole CTA              {(ph[:, 7] - ph[:, 1]).mean():9.0f}")
    t_first, t_last = ph[:, 1].min(), ph[:, 7].max()
    print(f"  kernel span {t_last - t_first} ns; sum of CTA lifetimes / (SMs x span) = "
          f"{(ph[:, 7] - ph[:, 1]).sum() / (148 * (t_last - t_first)):.2f}")
    for smid in (0, 1, 77):
        rows = ph[ph[:, 0] == smid]
        rows = rows[np.argsort(rows[:, 1])]
        print(f"  SM {smid}: " + " ".join(f"[{r[1] - t_first}..{r[7] - t_first}]" for r in rows))
t0 = tr[tr > 0].min()
names = {0: "softmax g0", 1: "softmax g1", 2: "mma for g0", 3: "mma for g1"}
for role in range(4):
    print(f"== {names[role]}  (cycles; stamps relative to the first)")
    for j in range(32):
        row = tr[role, j]
        if row[0] == 0:
            continue
        if role < 2:
            print(f"  j={j:2d} start {row[0] - t0:7d}  wait_S {row[1] - row[0]:5d}  softmax (ld, max, exp, pack, st) {row[4] - row[1]:5d}  "
                  f"st_wait+arrive {row[5] - row[4]:5d}  total {row[5] - row[0]:5d}")
        else:
            print(f"  j={j:2d} start {row[0] - t0:7d}  wait_P {row[1] - row[0]:5d}  issue PV+l {row[2] - row[1]:5d}  issue S(j+1) {row[3] - row[2]:5d}")
