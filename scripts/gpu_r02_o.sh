#!/usr/bin/env bash
# Round-2 GPU call O: attention kernels with whole-warp (uniform-register) MMA / TMA issue: timeline probe of both
# kernels, kernel + parity suites under each, bench A/B (pair vs pipe) on the same box.
set -u
mkdir -p gpurun_out
DV_ATTN_PIPE=2 timeout 200 python scripts/probe/attn_trace.py 3 2237 24 > gpurun_out/r02o_attn_trace_pair.txt 2>&1
DV_ATTN_PIPE=1 timeout 200 python scripts/probe/attn_trace.py 3 2237 24 > gpurun_out/r02o_attn_trace_pipe.txt 2>&1
grep -E "j= [345] " gpurun_out/r02o_attn_trace_pair.txt gpurun_out/r02o_attn_trace_pipe.txt
for m in 2 1; do
  DV_ATTN_PIPE=$m timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02o_pytest_$m.log 2>&1
  echo "pytest (DV_ATTN_PIPE=$m) rc=$?" | tee -a gpurun_out/r02o_summary.txt
  tail -1 gpurun_out/r02o_pytest_$m.log
done
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02o_prof_$name.csv > gpurun_out/r02o_bench_$name.json 2> gpurun_out/r02o_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02o_summary.txt
  python scripts/prof_table.py gpurun_out/r02o_prof_$name.csv 400 > gpurun_out/r02o_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02o_prof_$name.csv
}
run pair DV_ATTN_PIPE=2
run pipe1 DV_ATTN_PIPE=1
python - <<'PY' | tee -a gpurun_out/r02o_summary.txt
import json
for n in ("pair", "pipe1"):
    try:
        d = json.loads(open(f"gpurun_out/r02o_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
grep -E "^attn" gpurun_out/r02o_launch_table_pair.txt | head -12 | tee -a gpurun_out/r02o_summary.txt
grep -E "^attn" gpurun_out/r02o_launch_table_pipe1.txt | head -12 | tee -a gpurun_out/r02o_summary.txt
