#!/usr/bin/env bash
# Round-2 GPU call J: the persistent block kernel (LN / GEMM phases of a joint block in one launch): model suites with it,
# then bench A/B (rollout) against DV_MMDIT_PBK=0 on the same box.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mmdit or unit or cache" > gpurun_out/r02j_pytest_a.log 2>&1
echo "pytest parity rc=$?" | tee gpurun_out/r02j_summary.txt
tail -15 gpurun_out/r02j_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_boundary.py tests/test_gpu_rollout.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not vae and not conv" > gpurun_out/r02j_pytest_b.log 2>&1
echo "pytest rest rc=$?" | tee -a gpurun_out/r02j_summary.txt
tail -5 gpurun_out/r02j_pytest_b.log
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02j_prof_$name.csv > gpurun_out/r02j_bench_$name.json 2> gpurun_out/r02j_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02j_summary.txt
  python scripts/prof_table.py gpurun_out/r02j_prof_$name.csv 400 > gpurun_out/r02j_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02j_prof_$name.csv
}
run pbk DV_DUMMY=1
run nopbk DV_MMDIT_PBK=0
python - <<'PY' | tee -a gpurun_out/r02j_summary.txt
import json
for n in ("pbk", "nopbk"):
    try:
        d = json.loads(open(f"gpurun_out/r02j_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step launches {d['gpu_launches']} | " + " ".join(f"{k} {v['ms']:.0f}ms@{v['achieved']:.0f}" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
grep -E "^pbk" gpurun_out/r02j_launch_table_pbk.txt
