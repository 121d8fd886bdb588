#!/usr/bin/env bash
# Round-2 GPU call T: trimmed decode of continuation iterations (dv_vae_plan_set_first_frame): VAE / rollout / full-size
# suites (bit-identity of the kept frames), then the default bench with a launch table (compare with r02s).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_rollout.py tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/r02t_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02t_summary.txt
tail -4 gpurun_out/r02t_pytest_gpu.log | tee -a gpurun_out/r02t_summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-same-box-eager --profile-dump gpurun_out/r02t_prof.csv > gpurun_out/r02t_bench_rollout.json 2> gpurun_out/r02t_bench_rollout.err
echo "bench rc=$?" | tee -a gpurun_out/r02t_summary.txt
python scripts/prof_table.py gpurun_out/r02t_prof.csv 400 > gpurun_out/r02t_launch_table_rollout.txt 2>&1
rm -f gpurun_out/r02t_prof.csv
python - <<'PY' | tee -a gpurun_out/r02t_summary.txt
import json
d = json.loads(open("gpurun_out/r02t_bench_rollout.json").read().strip().splitlines()[-1])
c = d["roofline"]["classes"]
print(f"{d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step e2e {d['e2e']['value']:.2f} | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
print("clocks", d["clocks"])
PY
grep -E "^conv T(44|42|40|38|36|34|46|23|25|57|29) " gpurun_out/r02t_launch_table_rollout.txt | head -24 | tee -a gpurun_out/r02t_summary.txt
