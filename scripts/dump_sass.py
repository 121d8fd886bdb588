"""Write gzip'ed `cuobjdump -sass` listings of the hot kernels of the built library into profiles/sass/ plus an INDEX.txt
with the tcgen05 / TMA mnemonic counts per kernel.      python scripts/dump_sass.py
"""
import gzip
import re
import subprocess
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "deepv_b200" / "libdeepv_b200.so"
OUT = ROOT / "profiles" / "sass"
WANT = [("attn_pipe_kernel_k64", r"attn_pipe_kernelILi64E"), ("attn_pipe_kernel_k128", r"attn_pipe_kernelILi128E"),
        ("attn_pair_kernel", r"attn_pair_kernel"), ("attn_kernel", r"11attn_kernel"), ("conv_halo_kernel", r"conv_halo_kernel"),
        ("gemm_tc_kernel_mode1", r"gemm_tc_kernelILi1E"), ("gemm_tc_kernel_mode2", r"gemm_tc_kernelILi2E"),
        ("gemm_tc_kernel_mode4", r"gemm_tc_kernelILi4E"), ("gemm_tc_kernel_mode6", r"gemm_tc_kernelILi6E"),
        ("gemm_pair_kernel_mode1", r"gemm_pair_kernelILi1E"), ("gemm_pair_kernel_mode2", r"gemm_pair_kernelILi2E"),
        ("gemm_pair_kernel_mode4", r"gemm_pair_kernelILi4E"), ("gemm_pair_kernel_mode6", r"gemm_pair_kernelILi6E"),
        ("pbk_kernel", r"pbk_kernel"), ("gemv_kernel_2", r"gemv_kernelILi2E"), ("gn_apply_kernel", r"gn_apply_kernel"),
        ("ln_modulate_kernel_1536", r"ln_modulate_kernelILi1536E"), ("sp_barrier_kernel", r"sp_barrier_kernel")]
KEYS = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMACCTL", "LDTM", "STTM", "UTCBAR", "SYNCS", "ACQBULK", "MUFU.EX2",
        "MUFU.TANH", "FFMA2", "FADD2", "USETMAXREG", "ELECT", "BRA.U.ANY", "HMMA")

sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
funcs, name, buf = {}, None, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            funcs[name] = buf
        name, buf = m.group(1), []
    buf.append(line)
if name:
    funcs[name] = buf
OUT.mkdir(parents=True, exist_ok=True)
for old in OUT.glob("*.sass.gz"):
    old.unlink()
index = ["cuobjdump -sass listings (gzip) of the hot kernels of deepv_b200/libdeepv_b200.so, round 2 final build (scripts/dump_sass.py)",
         "(sm_100a; mnemonic counts: UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce, LDTM/STTM = tcgen05.ld/st,",
         " UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg, FFMA2/FADD2 = packed fp32x2; BRA.U.ANY = the ELECT / R2UR loop ptxas puts",
         " around a tcgen05 / TMA instruction issued from a divergent thread — the hot kernels issue from the whole warp; HMMA would be",
         " the legacy tensor-core path: none)", ""]
for label, pat in WANT:
    hit = [k for k in funcs if re.search(pat, k)]
    if not hit:
        index.append(f"{label:28s} NOT FOUND")
        continue
    lines = funcs[hit[0]]
    c = Counter()
    for ln in lines:
        for k in KEYS:
            if re.search(r"\b" + re.escape(k) + r"\b", ln):
                c[k] += 1
    with gzip.open(OUT / f"r02_{label}.sass.gz", "wt") as f:
        f.write("\n".join(lines))
    dem = subprocess.run(["c++filt", hit[0]], capture_output=True, text=True).stdout.strip()
    index.append(f"{label:28s} {len(lines):6d} lines  " + " ".join(f"{k}:{c[k]}" for k in KEYS if c[k]))
    index.append(f"    {dem[:150]}")
(OUT / "INDEX.txt").write_text("\n".join(index) + "\n")
print("\n".join(index))
