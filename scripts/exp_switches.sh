#!/usr/bin/env bash
# Validate and time the experimental switches left at the end of round 1 (run on a B200 box):
#   DV_ATTN_PIPE=1    double-buffered-S attention kernel   (attention.cu: attn_pipe_kernel)
#   DV_MMDIT_GRAPH=1  CUDA-graph replay of the MMDiT forward (mmdit.cu: forward_graph)
# For each: the model-level GPU suites, then the default bench; results under gpurun_out/switch_<name>.*
#   gpurun --timeout 1500 -- 'bash scripts/exp_switches.sh'
set -u
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/switch_baseline.json 2> gpurun_out/switch_baseline.err
for sw in DV_ATTN_PIPE DV_MMDIT_GRAPH; do
  env $sw=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_sharding.py \
      -m gpu -q -x > gpurun_out/switch_$sw.pytest.log 2>&1
  echo "$sw pytest rc=$?" | tee -a gpurun_out/switch_summary.txt
  env $sw=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/switch_$sw.json 2> gpurun_out/switch_$sw.err
  echo "$sw bench rc=$?" | tee -a gpurun_out/switch_summary.txt
done
python - <<'PY' | tee -a gpurun_out/switch_summary.txt
import json
for n in ("baseline", "DV_ATTN_PIPE", "DV_MMDIT_GRAPH"):
    try:
        d = json.load(open(f"gpurun_out/switch_{n}.json"))
        print(f"{n:16s} {d['value']:.2f} frames/s  {d['ms_per_step']:.1f} ms/step  sm_mhz {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "no result:", e)
PY
