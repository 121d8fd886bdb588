#!/usr/bin/env bash
# Round-2 GPU call R: dense problems with contiguous batch rows flattened into one row space (fewer partial tiles): kernel + parity
# + full-size suites, then the default bench with a launch table (compare with r02q on the previous build).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py -m gpu -x -q > gpurun_out/r02r_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02r_summary.txt
tail -2 gpurun_out/r02r_pytest_gpu.log | tee -a gpurun_out/r02r_summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-same-box-eager --profile-dump gpurun_out/r02r_prof.csv > gpurun_out/r02r_bench_rollout.json 2> gpurun_out/r02r_bench_rollout.err
echo "bench rc=$?" | tee -a gpurun_out/r02r_summary.txt
python scripts/prof_table.py gpurun_out/r02r_prof.csv 400 > gpurun_out/r02r_launch_table_rollout.txt 2>&1
rm -f gpurun_out/r02r_prof.csv
python - <<'PY' | tee -a gpurun_out/r02r_summary.txt
import json
d = json.loads(open("gpurun_out/r02r_bench_rollout.json").read().strip().splitlines()[-1])
c = d["roofline"]["classes"]
print(f"{d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step e2e {d['e2e']['value']:.2f} | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
print("clocks", d["clocks"])
PY
grep -E "^gemm" gpurun_out/r02r_launch_table_rollout.txt | head -14 | tee -a gpurun_out/r02r_summary.txt
