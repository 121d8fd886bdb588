#!/usr/bin/env bash
# Round-2 GPU call K: persistent block kernel, per-layout times and the row threshold (A/B on one box).
set -u
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02k_prof_$name.csv > gpurun_out/r02k_bench_$name.json 2> gpurun_out/r02k_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02k_summary.txt
  python scripts/prof_table.py gpurun_out/r02k_prof_$name.csv 600 > gpurun_out/r02k_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02k_prof_$name.csv
}
rm -f gpurun_out/r02k_summary.txt
run rows2048 DV_PBK_MAX_ROWS=2048
run rows1024 DV_PBK_MAX_ROWS=1024
run rows512 DV_PBK_MAX_ROWS=512
run nopbk DV_MMDIT_PBK=0
python - <<'PY' | tee -a gpurun_out/r02k_summary.txt
import json
for n in ("rows2048", "rows1024", "rows512", "nopbk"):
    try:
        d = json.loads(open(f"gpurun_out/r02k_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:10s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step launches {d['gpu_launches']} | " + " ".join(f"{k} {v['ms']:.0f}ms" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
grep -E "^pbk" gpurun_out/r02k_launch_table_rows2048.txt
