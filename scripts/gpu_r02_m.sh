#!/usr/bin/env bash
# Round-2 GPU call M: TMA L2 prefetch of the weight panels of small dense problems: PBK phase trace, GEMM kernel tests,
# rollout bench A/B against DV_GEMM_NO_L2_PREFETCH=1 on the same box.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02m_summary.txt
timeout 300 python scripts/probe/pbk_trace.py > gpurun_out/r02m_pbk_trace.txt 2>&1
cat gpurun_out/r02m_pbk_trace.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02m_summary.txt
tail -2 gpurun_out/r02m_pytest.log
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-same-box-eager \
      --profile-dump gpurun_out/r02m_prof_$name.csv > gpurun_out/r02m_bench_$name.json 2> gpurun_out/r02m_bench_$name.err
  echo "bench $name rc=$?" | tee -a gpurun_out/r02m_summary.txt
  python scripts/prof_table.py gpurun_out/r02m_prof_$name.csv 600 > gpurun_out/r02m_launch_table_$name.txt 2>&1
  rm -f gpurun_out/r02m_prof_$name.csv
}
run prefetch DV_DUMMY=1
run noprefetch DV_GEMM_NO_L2_PREFETCH=1
run prefetch_pbk512 DV_MMDIT_PBK=1 DV_PBK_MAX_ROWS=512
python - <<'PY' | tee -a gpurun_out/r02m_summary.txt
import json
for n in ("prefetch", "noprefetch", "prefetch_pbk512"):
    try:
        d = json.loads(open(f"gpurun_out/r02m_bench_{n}.json").read().strip().splitlines()[-1])
        c = d["roofline"]["classes"]
        print(f"{n:16s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step launches {d['gpu_launches']} | " + " ".join(f"{k} {v['ms']:.0f}ms" for k, v in c.items()))
    except Exception as e:
        print(n, "no result:", e)
PY
for n in prefetch noprefetch; do echo "== $n"; grep -E "^gemm B2 M96\+77|^gemm B2 M77 |^gemm B3 M269|^gemm B2 M384\+77 N1536" gpurun_out/r02m_launch_table_$n.txt; done
