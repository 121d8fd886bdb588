#!/usr/bin/env bash
# Round-2 GPU call U (final build of the round) (N GPUs, N = $1): sharded-vs-alone parity of the rollout group, then the rollout bench as ONE
# rollout over all N GPUs (CFG branch groups x Ulysses over peer memory, VAE tiles / encodes dealt over the ranks).
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR scripts/check_sharded.py --group $N > gpurun_out/r02u_check_sharded_${N}gpu.log 2>&1
echo "check_sharded $N rc=$?" | tee gpurun_out/r02u_summary_${N}.txt
grep -E "branches|decode|SHARDED" gpurun_out/r02u_check_sharded_${N}gpu.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r02u_bench_${N}gpu.json 2> gpurun_out/r02u_bench_${N}gpu.err
echo "bench $N rc=$?" | tee -a gpurun_out/r02u_summary_${N}.txt
python - <<PY | tee -a gpurun_out/r02u_summary_${N}.txt
import json
try:
    d = json.loads(open("gpurun_out/r02u_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print(f"N=${N}: {d['value']:.2f} frames/s {d['ms_per_step']:.1f} ms/step launches {d['gpu_launches']} e2e {d.get('e2e') and round(d['e2e']['value'],2)} replicas {d.get('replicas') and round(d['replicas']['value'],2)}")
    print(d['config']['parallelism'])
    print({k: (round(v['ms']), v['achieved']) for k, v in d['roofline']['classes'].items()})
except Exception as e:
    print("no result:", e)
PY
tail -5 gpurun_out/r02u_bench_${N}gpu.err
