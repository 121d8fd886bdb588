import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deepv_b200 import _lib
lib = _lib.load()
raw = C.CDLL(str(Path(__file__).resolve().parents[1] / "deepv_b200" / "libdeepv_b200.so"))
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
B, L, H = 2, 1613, 24
qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
sizes = [77, 768, 768]
bounds, acc = [], 0
for s in sizes:
    acc += s; bounds.append(acc)
kv = torch.empty(L, dtype=torch.int32); pos = 0
for i, s in enumerate(sizes):
    kv[pos:pos + s] = bounds[max(i, 1)]; pos += s
Lpad = (L + 127) // 128 * 128
kb = torch.zeros(B, Lpad, device="cuda"); kb[:, L:] = float("-inf"); kb[0, 1:77] = float("-inf")
out = torch.empty(B, L, H * 64, device="cuda", dtype=torch.bfloat16)
kvd = kv.cuda()
run = lambda: _lib.check(lib.dv_attention(p(qkv), p(out), p(kvd), p(kb), B, L, Lpad, H, None))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"attn B{B} L{L} H{H}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
