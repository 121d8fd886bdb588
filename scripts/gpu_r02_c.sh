#!/usr/bin/env bash
# Round-2 GPU call C (2 GPUs): Ulysses over peer memory with the device barrier + graph replay: virtual-rank tests,
# sharded-vs-alone parity under torchrun, then the rollout bench as ONE rollout over both GPUs, A/B of the switches.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -x -q > gpurun_out/r02c_pytest_sharding.log 2>&1
echo "pytest sharding rc=$?" | tee gpurun_out/r02c_summary.txt
tail -3 gpurun_out/r02c_pytest_sharding.log
timeout 600 $TR scripts/check_sharded.py --group 2 > gpurun_out/r02c_check_sharded_2gpu.log 2>&1
echo "check_sharded rc=$?" | tee -a gpurun_out/r02c_summary.txt
grep -E "branches|decode|SHARDED" gpurun_out/r02c_check_sharded_2gpu.log
timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err
echo "bench 2gpu rc=$?" | tee -a gpurun_out/r02c_summary.txt
DV_MMDIT_GRAPH=0 timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-replicas > gpurun_out/r02c_bench_2gpu_nograph.json 2> gpurun_out/r02c_bench_2gpu_nograph.err
echo "bench 2gpu nograph rc=$?" | tee -a gpurun_out/r02c_summary.txt
DV_MMDIT_GRAPH=0 DV_SP_NCCL_BARRIER=1 timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-replicas > gpurun_out/r02c_bench_2gpu_ncclbar.json 2> gpurun_out/r02c_bench_2gpu_ncclbar.err
echo "bench 2gpu nccl barrier rc=$?" | tee -a gpurun_out/r02c_summary.txt
python - <<'PY' | tee -a gpurun_out/r02c_summary.txt
import json
for n in ("2gpu", "2gpu_nograph", "2gpu_ncclbar"):
    try:
        d = json.loads(open(f"gpurun_out/r02c_bench_{n}.json").read().strip().splitlines()[-1])
        print(f"{n:14s} {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step launches {d['gpu_launches']} e2e {d.get('e2e') and d['e2e']['value']} replicas {d.get('replicas') and d['replicas']['value']}")
    except Exception as e:
        print(n, "no result:", e)
PY
tail -5 gpurun_out/r02c_bench_2gpu.err
