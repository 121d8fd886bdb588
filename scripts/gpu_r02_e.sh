#!/usr/bin/env bash
# Round-2 GPU call E: conditioning cache + full-size first-iteration parity test; full bench line; A/B of the cache.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02e_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02e_summary.txt
grep -E "full-size iteration|passed|failed|Error" gpurun_out/r02e_pytest_gpu.log | tail -8
timeout 1200 python bench.py --steps 3 --warmup 2 --profile-dump gpurun_out/r02e_prof.csv > gpurun_out/r02e_bench_rollout.json 2> gpurun_out/r02e_bench_rollout.err
echo "bench rc=$?" | tee -a gpurun_out/r02e_summary.txt
python scripts/prof_table.py gpurun_out/r02e_prof.csv 400 > gpurun_out/r02e_launch_table_rollout.txt 2>&1
gzip -f gpurun_out/r02e_prof.csv
DV_MOD_CACHE_SLOTS=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager > gpurun_out/r02e_bench_nocache.json 2> gpurun_out/r02e_bench_nocache.err
echo "bench nocache rc=$?" | tee -a gpurun_out/r02e_summary.txt
python - <<'PY' | tee -a gpurun_out/r02e_summary.txt
import json
for n in ("rollout", "nocache"):
    try:
        d = json.loads(open(f"gpurun_out/r02e_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step e2e {d.get('e2e') and round(d['e2e']['value'],2)} | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
