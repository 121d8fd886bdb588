#!/usr/bin/env bash
# Round-2 GPU call N: attention with two query tiles per CTA (attn_pair_kernel, setmaxnreg-partitioned register file):
# kernel + model parity suites with DV_ATTN_PIPE=2, then bench A/B against the one-group pipelined kernel on the same box.
set -u
mkdir -p gpurun_out
DV_ATTN_PIPE=2 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1
echo "pytest (pair) rc=$?" | tee gpurun_out/r02n_summary.txt
tail -3 gpurun_out/r02n_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02n_prof_$name.csv > gpurun_out/r02n_bench_$name.json 2> gpurun_out/r02n_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02n_summary.txt
  python scripts/prof_table.py gpurun_out/r02n_prof_$name.csv 400 > gpurun_out/r02n_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02n_prof_$name.csv
}
run pair DV_ATTN_PIPE=2
run pipe1 DV_ATTN_PIPE=1
python - <<'PY' | tee -a gpurun_out/r02n_summary.txt
import json
for n in ("pair", "pipe1"):
    try:
        d = json.loads(open(f"gpurun_out/r02n_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
grep -E "^attn" gpurun_out/r02n_launch_table_pair.txt | head -12 | tee -a gpurun_out/r02n_summary.txt
grep -E "^attn" gpurun_out/r02n_launch_table_pipe1.txt | head -12 | tee -a gpurun_out/r02n_summary.txt
