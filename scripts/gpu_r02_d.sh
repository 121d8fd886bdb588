#!/usr/bin/env bash
# Round-2 GPU call D: TMA-store / TMA-reduce epilogues of the dense GEMMs + the "no split-K with >= 148 tiles" rule:
# the whole GPU suite, then the rollout bench A/B against each switch on the same box.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/r02d_summary.txt
tail -3 gpurun_out/r02d_pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02d_prof_$name.csv > gpurun_out/r02d_bench_$name.json 2> gpurun_out/r02d_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02d_summary.txt
  python scripts/prof_table.py gpurun_out/r02d_prof_$name.csv 400 > gpurun_out/r02d_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02d_prof_$name.csv
}
run default DV_DUMMY=1
run no_tma_store DV_GEMM_NO_TMA_STORE=1
run split_large DV_GEMM_SPLIT_LARGE=1
python - <<'PY' | tee -a gpurun_out/r02d_summary.txt
import json
for n in ("default", "no_tma_store", "split_large"):
    try:
        d = json.loads(open(f"gpurun_out/r02d_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:14s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
