#!/usr/bin/env bash
# Round-2 GPU call V: the final build — whole GPU suite, smoke(), both bench arms as the driver runs them, the C2 unit
# workload, an ncu gpu__time_duration launch list of one unit step, and `ncu --set full` of one launch per kernel class.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02v_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02v_summary.txt
grep -E "passed|failed" gpurun_out/r02v_pytest_gpu.log | tail -2 | tee -a gpurun_out/r02v_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02v_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/r02v_summary.txt
tail -1 gpurun_out/r02v_smoke.log | tee -a gpurun_out/r02v_summary.txt
timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 --profile-dump gpurun_out/r02v_prof.csv > gpurun_out/r02v_bench_rollout.json 2> gpurun_out/r02v_bench_rollout.err
echo "bench rc=$?" | tee -a gpurun_out/r02v_summary.txt
python scripts/prof_table.py gpurun_out/r02v_prof.csv 400 > gpurun_out/r02v_launch_table_rollout.txt 2>&1
gzip -f gpurun_out/r02v_prof.csv
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/r02v_bench_reference.json 2> gpurun_out/r02v_bench_reference.err
echo "bench reference rc=$?" | tee -a gpurun_out/r02v_summary.txt
timeout 600 python bench.py --workload unit --steps 5 --warmup 3 --no-cpu-baseline --no-same-box-eager --profile-dump gpurun_out/r02v_prof_unit.csv > gpurun_out/r02v_bench_unit.json 2> gpurun_out/r02v_bench_unit.err
echo "bench unit rc=$?" | tee -a gpurun_out/r02v_summary.txt
python scripts/prof_table.py gpurun_out/r02v_prof_unit.csv 400 > gpurun_out/r02v_launch_table_unit.txt 2>&1
rm -f gpurun_out/r02v_prof_unit.csv
# launch list (one C2 unit step; the profiler serialises kernels and flushes caches: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02v_ncu_launches_unit.csv \
    python bench.py --workload unit --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-same-box-eager > gpurun_out/r02v_ncu_unit.log 2>&1
echo "ncu launch list rc=$?" | tee -a gpurun_out/r02v_summary.txt
wc -l gpurun_out/r02v_ncu_launches_unit.csv
gzip -f gpurun_out/r02v_ncu_launches_unit.csv
python scripts/ncu_target.py all > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|gemm_|attn_" -f -o gpurun_out/r02v_ncu_targets \
    python scripts/ncu_target.py all > gpurun_out/r02v_ncu.log 2>&1
echo "ncu targets rc=$?" | tee -a gpurun_out/r02v_summary.txt
python - <<'PY' | tee -a gpurun_out/r02v_summary.txt
import json
for n in ("rollout", "reference", "unit"):
    try:
        d = json.loads(open(f"gpurun_out/r02v_bench_{n}.json").read().strip().splitlines()[-1])
        print(f"{n:10s} {d['value']:.3f} frames/s {d['ms_per_step']:.1f} ms/step e2e {d.get('e2e') and round(d['e2e']['value'],3)} eager {d.get('same_box_eager') and d['same_box_eager'].get('value')} cpu {d.get('cpu_baseline') and d['cpu_baseline']['value']}")
    except Exception as e:
        print(n, "no result:", e)
PY
