"""Probe (GPU): time the attention kernel selected by DV_ATTN_PIPE on the token layouts of a rollout (deepv_b200.work).
    DV_ATTN_PIPE=3 python scripts/probe/attn_time.py
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from deepv_b200 import _lib, work  # noqa: E402

lib = _lib.load()
H = 24
cases = []
fw = work.rollout_forwards(2)
seen = set()
for f in fw:
    key = (f["B"], tuple(f["clips"]), f["hist"])
    if key in seen:
        continue
    seen.add(key)
    cases.append(f)
pick = [c for c in cases if c["stage"] == 2][-1], [c for c in cases if c["stage"] == 2 and c["B"] == 2][-1], \
    [c for c in cases if c["stage"] == 1][-1], [c for c in cases if c["stage"] == 0][-1]
tot = 0.0
for f in pick:
    B = f["B"]
    lv, lc = work.mmdit_tokens(f["clips"], f["hist"])
    L = lv + lc
    frames = []
    for t, h, w in f["clips"]:
        frames += [(h // 2) * (w // 2)] * t
    kv = torch.empty(L, dtype=torch.int32)
    first_end = lc + frames[0]
    kv[:lc] = first_end
    pos = lc
    for n in frames:
        kv[pos:pos + n] = pos + n
        pos += n
    Lpad = (L + 127) // 128 * 128
    kb = torch.zeros(B, Lpad, device="cuda")
    kb[:, L:] = float("-inf")
    kb[:, 30:77] = float("-inf")          # padded text tokens
    qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
    out = torch.empty(B, L, H * 64, device="cuda", dtype=torch.bfloat16)
    kvd = kv.cuda()
    args = (qkv.data_ptr(), out.data_ptr(), kvd.data_ptr(), kb.data_ptr(), B, L, Lpad, H, None)
    for _ in range(5):
        _lib.check(lib.dv_attention(*args))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        lib.dv_attention(*args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / n
    fl = work.mmdit_flops(B, f["clips"], f["hist"])["attention"] / work.N_LAYERS
    tot += us
    print(f"mode {os.environ.get('DV_ATTN_PIPE', 'default')}  B{B} L{L} stage {f['stage']}: {us:7.1f} us  {fl / us / 1e6:6.1f} TFLOP/s (visible pairs)", flush=True)
print(f"mode {os.environ.get('DV_ATTN_PIPE', 'default')}  sum {tot:.1f} us")
