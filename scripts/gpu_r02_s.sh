#!/usr/bin/env bash
# Round-2 GPU call S: (1) split-K reduction with the peers' loads scheduled together (gemm_trace probe: reduce / epilogue
# phase was 5.4 us), (2) GroupNorm-apply A/B: SiLU through one MUFU.TANH instead of EX2 + RCP, pixels per CTA.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/probe/gemm_trace.py --rebuild 2>&1 | grep -v "^$" | head -48 > gpurun_out/r02s_gemm_trace.txt
grep -E "^==|reduce / epilogue|accumulator->all" gpurun_out/r02s_gemm_trace.txt | head -12
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r02s_summary.txt
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02s_prof_$name.csv > gpurun_out/r02s_bench_$name.json 2> gpurun_out/r02s_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02s_summary.txt
  python scripts/prof_table.py gpurun_out/r02s_prof_$name.csv 400 > gpurun_out/r02s_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02s_prof_$name.csv
}
run tanh DV_DUMMY=1
run exp DV_GN_SILU_EXP=1
run tanh_ppc64 DV_GN_PPC=64
python - <<'PY' | tee -a gpurun_out/r02s_summary.txt
import json
for n in ("tanh", "exp", "tanh_ppc64"):
    try:
        d = json.loads(open(f"gpurun_out/r02s_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
for n in tanh exp tanh_ppc64; do grep -E "^gn_apply" gpurun_out/r02s_launch_table_$n.txt | sed "s/^/$n  /" | tee -a gpurun_out/r02s_summary.txt; done
grep -E " s[248] e2" gpurun_out/r02s_launch_table_tanh.txt | head -5 | tee -a gpurun_out/r02s_summary.txt
