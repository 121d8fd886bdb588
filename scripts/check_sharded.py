"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): one AR unit + the RGB/disparity decodes
computed by a rollout group (CFG branch groups x Ulysses sequence parallelism, VAE tiles dealt over
the ranks) must match the same computation done by every rank alone.

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 scripts/check_sharded.py --group 4
"""
import argparse
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deepv_b200 import synthetic as synth  # noqa: E402
from deepv_b200.mmdit import B200MMDiT  # noqa: E402
from deepv_b200.parallel import Shard  # noqa: E402
from deepv_b200.pipeline import B200Pipeline  # noqa: E402
from deepv_b200.scheduler import B200Scheduler  # noqa: E402
from deepv_b200.vae import B200VAE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--group", type=int, default=0)
ap.add_argument("--layers", type=int, default=2)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
gsz = args.group or world
shard = Shard.grouped(gsz)
shard.setup_sp(dev)

cfg, W = synth.mmdit_weights(dict(num_layers=args.layers), seed=1)
over = dict(decoder_block_out_channels=(128, 128, 128, 128), encoder_block_out_channels=(128, 128, 128, 128),
            decoder_layers_per_block=(1, 1, 1, 1))
vcfg, VW = synth.vae_weights(over, seed=2)
sched = dict(num_train_timesteps=1000, shift=1.0, stages=3, stage_range=[0, 1 / 3, 2 / 3, 1], gamma=0.3333)


def make_pipe():
    dit = B200MMDiT(W, cfg, device=dev)
    vae = B200VAE(VW, vcfg, device=dev, dtype=torch.bfloat16)
    vae.enable_tiling()
    return B200Pipeline(dit, vae, B200Scheduler(**sched), device=dev, torch_dtype=torch.bfloat16)


ok = True
for n_branch, hist in ((2, False), (3, True)):
    g = torch.Generator().manual_seed(7 + n_branch)
    bf = torch.bfloat16
    lat = (torch.randn(1, 38, 1, 12, 16, generator=g) * 2).to(bf).to(dev)
    conds = [[torch.randn(n_branch, 38, 1, 12 * 2 ** s, 16 * 2 ** s, generator=g).to(bf).to(dev)] for s in range(3)]
    noise = [torch.randn(1, 38, 1, 24, 32, generator=g).to(bf).to(dev), torch.randn(1, 38, 1, 48, 64, generator=g).to(bf).to(dev)]
    enc = torch.randn(n_branch, 77, 4096, generator=g).to(bf).to(dev)
    pooled = torch.randn(n_branch, 2048, generator=g).to(dev)
    mask = torch.zeros(n_branch, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1:, :12] = 1
    mask = mask.to(dev)
    history = torch.randn(1, 38, 1, 48, 64, generator=g).to(bf).to(dev) if hist else None
    steps = [2, 2, 2]
    alone = make_pipe().generate_one_unit(lat, history, conds, enc, mask, pooled, steps, block_noise=noise)
    group = make_pipe().generate_one_unit(lat, history, conds, enc, mask, pooled, steps, block_noise=noise, shard=shard)
    torch.cuda.synchronize()
    for s, (a, b) in enumerate(zip(alone, group)):
        err = ((a.float() - b.float()).abs().max() / a.float().abs().max()).item()
        layout = shard.layout(n_branch)
        if rank == 0:
            print(f"branches {n_branch} (groups x sp = {layout[0]} x {layout[1]}) stage {s}: max rel diff {err:.3e}", flush=True)
        ok &= err <= 2e-2
z = [torch.randn(1, 16, 2, 48, 64, generator=torch.Generator().manual_seed(5 + i)).to(torch.bfloat16).to(dev) for i in range(2)]
p1, p2 = make_pipe(), make_pipe()
a = [p1.decode_latent(t) for t in z]
b = p2.decode_latents_sharded(z, shard)
torch.cuda.synchronize()
for i in range(2):
    same = torch.equal(a[i], b[i])
    if rank == 0:
        print(f"decode {i}: tiles over {shard.world} ranks bit-equal to the single-rank decode: {same}", flush=True)
    ok &= same
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED PARITY", "OK" if flag.item() == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
