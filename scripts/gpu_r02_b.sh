#!/usr/bin/env bash
# Round-2 GPU call B: the new bench (rollout workload, both arms, per-class roofline, same-box eager), the C2 unit
# workload for continuity with round 1, and `ncu --set full` captures of one launch of each hot kernel.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 2 --profile-dump gpurun_out/r02b_prof_rollout.csv > gpurun_out/r02b_bench_rollout.json 2> gpurun_out/r02b_bench_rollout.err
echo "bench rollout rc=$?" | tee gpurun_out/r02b_summary.txt
python scripts/prof_table.py gpurun_out/r02b_prof_rollout.csv 400 > gpurun_out/r02b_launch_table_rollout.txt 2>&1
gzip -f gpurun_out/r02b_prof_rollout.csv
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
echo "bench reference rc=$?" | tee -a gpurun_out/r02b_summary.txt
timeout 600 python bench.py --workload unit --steps 5 --warmup 3 --no-cpu-baseline --no-same-box-eager --profile-dump gpurun_out/r02b_prof_unit.csv > gpurun_out/r02b_bench_unit.json 2> gpurun_out/r02b_bench_unit.err
echo "bench unit rc=$?" | tee -a gpurun_out/r02b_summary.txt
python scripts/prof_table.py gpurun_out/r02b_prof_unit.csv 400 > gpurun_out/r02b_launch_table_unit.txt 2>&1
rm -f gpurun_out/r02b_prof_unit.csv
python scripts/ncu_target.py all > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|gemm_|attn_" -f -o gpurun_out/r02b_ncu_targets \
    python scripts/ncu_target.py all > gpurun_out/r02b_ncu.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r02b_summary.txt
ls -la gpurun_out/ | tail -20
