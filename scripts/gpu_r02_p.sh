#!/usr/bin/env bash
# Round-2 GPU call P: full GPU suite + smoke + default bench on the build whose default attention kernel is the 64-key,
# two-CTAs-per-SM pipelined kernel (whole-warp MMA / TMA issue, FFMA2 / polynomial exp2, row sums in registers).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/r02p_summary.txt
tail -3 gpurun_out/r02p_pytest_gpu.log | tee -a gpurun_out/r02p_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/r02p_summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --profile-dump gpurun_out/r02p_prof.csv > gpurun_out/r02p_bench_rollout.json 2> gpurun_out/r02p_bench_rollout.err
echo "bench rc=$?" | tee -a gpurun_out/r02p_summary.txt
python scripts/prof_table.py gpurun_out/r02p_prof.csv 400 > gpurun_out/r02p_launch_table_rollout.txt 2>&1
gzip -f gpurun_out/r02p_prof.csv
python - <<'PY' | tee -a gpurun_out/r02p_summary.txt
import json
d = json.loads(open("gpurun_out/r02p_bench_rollout.json").read().strip().splitlines()[-1])
c = d["roofline"]["classes"]
print(f"{d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step e2e {d['e2e']['value']:.2f} | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
print("clocks", d["clocks"])
PY
grep -E "^attn" gpurun_out/r02p_launch_table_rollout.txt | head -8 | tee -a gpurun_out/r02p_summary.txt
