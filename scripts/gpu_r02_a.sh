#!/usr/bin/env bash
# Round-2 GPU call A: the whole GPU suite (incl. production-size parity + boundary tests), then the
# experimental switches (attention pipe kernel, CUDA-graph forward) validated and timed A/B on one box.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02a_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02a_summary.txt
tail -3 gpurun_out/r02a_pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_baseline.json 2> gpurun_out/r02a_baseline.err
for sw in DV_ATTN_PIPE DV_MMDIT_GRAPH; do
  env $sw=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_sharding.py \
      tests/test_gpu_fullsize.py tests/test_gpu_boundary.py -k "not conv and not vae" -m gpu -q -x > gpurun_out/r02a_$sw.pytest.log 2>&1
  echo "$sw pytest rc=$?" | tee -a gpurun_out/r02a_summary.txt
  env $sw=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_$sw.json 2> gpurun_out/r02a_$sw.err
  echo "$sw bench rc=$?" | tee -a gpurun_out/r02a_summary.txt
done
env DV_ATTN_PIPE=1 DV_MMDIT_GRAPH=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02a_both.json 2> gpurun_out/r02a_both.err
python - <<'PY' | tee -a gpurun_out/r02a_summary.txt
import json
for n in ("baseline", "DV_ATTN_PIPE", "DV_MMDIT_GRAPH", "both"):
    try:
        d = json.load(open(f"gpurun_out/r02a_{n}.json"))
        k = d["roofline"]["by_kind"]
        print(f"{n:16s} {d['value']:.2f} frames/s  {d['ms_per_step']:.1f} ms/step  sm_mhz {d['clocks']['sm_mhz']}  launches {d['gpu_launches']}"
              f"  dense {k['gemm_dense']['ms']:.1f} conv {k['gemm_conv']['ms']:.1f} attn {k['attention']['ms']:.1f} ms")
    except Exception as e:
        print(n, "no result:", e)
PY
