#!/usr/bin/env bash
# Round-2 GPU call H: attention with two softmax groups (attn_pipe2_kernel): kernel + model suites, then bench A/B
# against the one-group pipelined kernel (DV_ATTN_PIPE=1) on the same box.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_sharding.py tests/test_gpu_boundary.py tests/test_gpu_rollout.py -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02h_summary.txt
tail -2 gpurun_out/r02h_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02h_prof_$name.csv > gpurun_out/r02h_bench_$name.json 2> gpurun_out/r02h_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02h_summary.txt
  python scripts/prof_table.py gpurun_out/r02h_prof_$name.csv 400 > gpurun_out/r02h_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02h_prof_$name.csv
}
run pipe2 DV_DUMMY=1
run pipe1 DV_ATTN_PIPE=1
python - <<'PY' | tee -a gpurun_out/r02h_summary.txt
import json
for n in ("pipe2", "pipe1"):
    try:
        d = json.loads(open(f"gpurun_out/r02h_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
grep -E "^attn" gpurun_out/r02h_launch_table_pipe2.txt | head -8
grep -E "^attn" gpurun_out/r02h_launch_table_pipe1.txt | head -8
