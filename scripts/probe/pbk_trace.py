"""Probe (GPU): where does the time of one persistent-block-kernel launch go?  DV_PBK_TRACE=1 makes every CTA stamp
globaltimer at the end of each phase; this prints, per phase, when the last CTA finished its work and how long the
barrier behind it took (first-unit layouts of stage 0 / stage 1, 2-block model)."""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["DV_MMDIT_PBK"] = "1"
os.environ["DV_PBK_MAX_ROWS"] = "4096"
os.environ["DV_PBK_TRACE"] = os.environ.get("DV_PBK_TRACE", "1")
os.environ["DV_MOD_CACHE_SLOTS"] = "0"
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from deepv_b200 import _lib, synthetic as synth
from deepv_b200.mmdit import B200MMDiT

cfg, W = synth.mmdit_weights(dict(num_layers=3), seed=1)
model = B200MMDiT(W, cfg, out_dtype=torch.float32)
lib = model.lib
ptr, nbytes = C.c_void_p(), C.c_longlong()
_lib.check(lib.dv_mmdit_debug_buffer(model._handle, 0, C.byref(ptr), C.byref(nbytes)))
g = torch.Generator().manual_seed(0)
for name, dims in (("stage 0, B2 96+77", [(1, 12, 16), (1, 12, 16)]), ("stage 1, B2 384+77", [(1, 24, 32), (1, 24, 32)])):
    clips = [torch.randn(2, 38, *d, generator=g).cuda() for d in dims]
    enc, pooled = torch.randn(2, 77, 4096, generator=g).cuda(), torch.randn(2, 2048, generator=g).cuda()
    mask = torch.zeros(2, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1, :12] = 1
    for rep in range(3):
        model(sample=[clips], timestep_ratio=torch.full((2,), 500.0).cuda(), encoder_hidden_states=enc,
              encoder_attention_mask=mask.cuda(), pooled_projections=pooled)
    torch.cuda.synchronize()
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    rows = 13
    host = (C.c_ulonglong * (rows * n_sm))()
    cudart = C.CDLL("libcudart.so.12")
    assert cudart.cudaMemcpy(host, C.c_void_p(ptr.value + 64), C.c_size_t(rows * n_sm * 8), 2) == 0
    st = torch.tensor(list(host), dtype=torch.int64).view(rows, n_sm)
    t0 = st[0].min().item()
    print(f"== {name}: the LAST block launch (out, LN, FF1, FF2 of the last block); ns since the first CTA started")
    prev_end = st[0].max().item()
    for ph in range(1, rows):
        if st[ph].max().item() <= t0:
            break
        first, last = st[ph].min().item(), st[ph].max().item()
        print(f"  phase {ph - 1}: first CTA done at {first - t0:7d}  last at {last - t0:7d}   (phase span {last - prev_end:6d} ns)")
        prev_end = last
