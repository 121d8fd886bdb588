"""Whole-rollout timing of `B200Rollout.generate` (rows f1-f4 around the hot path) at the reference's
run.py configuration: 384x512, 24-block MMDiT, full VAE, 3 stages x 5 steps, 8 units per iteration.

    python scripts/bench_rollout.py [--iters 2] [--layers 24] [--cpu-feedback]

Prints one JSON line: emitted frames / s over the whole rollout (first iteration 57 frames, 32 per
further iteration), per-iteration generate_i2v and feedback times from CUDA events on the launching
stream, and (with --cpu-feedback) the oracle's host arithmetic for the same feedback step, VAE
history encode excluded, as the CPU figure beside it.  Synthetic seeded weights and a synthetic frame.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--cpu-feedback", action="store_true")
    ap.add_argument("--profile-dump", default=None, help="per-launch CSV of one extra profiled rollout")
    ap.add_argument("--check", action="store_true", help="under torchrun: also run the rollout un-sharded on "
                                                         "every rank and compare")
    args = ap.parse_args()
    out_fd = os.dup(1)                 # keep the JSON line alone on stdout: NCCL / library chatter goes to stderr
    os.dup2(2, 1)
    import torch
    from deepv_b200 import _lib, synthetic as synth
    from deepv_b200.mmdit import B200MMDiT
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.rollout import B200Rollout, DeviceNoise, PromptCache
    from deepv_b200.scheduler import B200Scheduler
    from deepv_b200.vae import B200VAE
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shard = None
    dist = None
    if world > 1:                      # one rollout group over all ranks: CFG branches x Ulysses, VAE tiles dealt out
        import torch.distributed as dist
        from deepv_b200.parallel import Shard
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        shard = Shard.grouped(world)
        shard.setup_sp(dev)
    lib = _lib.load()
    dtype = torch.bfloat16
    cfg, W = synth.mmdit_weights(dict(num_layers=args.layers), seed=1)
    dit = B200MMDiT(W, cfg, device=dev)
    del W
    vcfg, VW = synth.vae_weights(None, seed=2, encoder=True)
    vae = B200VAE(VW, vcfg, device=dev, dtype=dtype)
    vae.enable_tiling()
    del VW
    pipe = B200Pipeline(dit, vae, B200Scheduler(num_train_timesteps=1000, shift=1.0, stages=3,
                                                stage_range=[0, 1 / 3, 2 / 3, 1], gamma=0.3333),
                        model_cfg=dict(num_inference_steps=args.steps), device=dev, torch_dtype=dtype)
    g = torch.Generator().manual_seed(3)
    table = {k: dict(prompt_embeds=torch.randn(1, 77, 4096, generator=g).to(dtype),
                     pooled_prompt_embeds=torch.randn(1, 2048, generator=g),
                     prompt_attention_mask=(torch.arange(77) < n).long().view(1, 77))
             for k, n in (("empty", 1), ("w", 12), ("a", 12), ("d", 12))}
    ro = B200Rollout(pipe, PromptCache(table, None, dev))
    img = (torch.rand(384, 512, 3, generator=g) * 255).to(torch.uint8)
    n_prompts = 8 + 4 * (args.iters - 1)
    batch = dict(img=img, prompt=(["w", "a", "d", "w"] * n_prompts)[:n_prompts], prompt_type="action")

    best = None
    for rep in range(args.repeats + 1):                      # the first pass builds every plan (warm-up)
        lib.dv_launch_count_reset()
        events = []
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = ro.generate(batch, events=events, shard=shard,
                          noise=DeviceNoise(pipe, torch.Generator(device=dev).manual_seed(1234)))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        launches = lib.dv_launch_count()
        i2v = [e[0].elapsed_time(e[1]) for e in events]
        fb = [e[1].elapsed_time(e[2]) for e in events]
        total_ms = events[0][0].elapsed_time(events[-1][2])
        if world > 1:                                        # device time, max over ranks
            t = torch.tensor([total_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = t.item()
        frames = res["pred_img"].shape[2]
        rec = dict(frames=frames, total_ms=total_ms, wall_ms=wall * 1e3, i2v_ms=i2v, feedback_ms=fb, launches=launches)
        if rank == 0:
            print(("warm-up " if rep == 0 else "timed   ") + json.dumps(rec), file=sys.stderr, flush=True)
        if rep > 0 and (best is None or total_ms < best["total_ms"]):
            best = rec
    assert torch.isfinite(res["pred_img"]).all()
    if args.profile_dump:
        lib.dv_profile_enable(1)
        lib.dv_profile_reset()
        ro.generate(batch)
        torch.cuda.synchronize()
        lib.dv_profile_enable(0)
        _lib.check(lib.dv_profile_dump(args.profile_dump.encode()), "dv_profile_dump")
        lib.dv_profile_reset()
    check = None
    if args.check and shard is not None:
        trace = []
        alone = ro.generate(batch, noise=DeviceNoise(pipe, torch.Generator(device=dev).manual_seed(1234)), trace=trace)
        t1 = trace[1]                  # the second iteration again, from identical inputs, un-sharded and sharded
        forced = []
        for sh in (None, shard):
            forced.append(ro.generate_i2v(t1["motion_prompt"], True, t1["frames"], t1["input_disparity"], t1["input_raymap"],
                                          t1["input_history"], temp=8, num_inference_steps=args.steps, shard=sh,
                                          noise=DeviceNoise(pipe, torch.Generator(device=dev).manual_seed(99))))
        torch.cuda.synchronize()
        import math
        def psnr(a, b):
            return 10 * math.log10(4.0 / max(((a.float() - b.float()) ** 2).mean().item(), 1e-20))
        check = {"psnr_first_iteration_db": psnr(res["pred_img"][:, :, :57], alone["pred_img"][:, :, :57]),
                 "psnr_all_frames_db": psnr(res["pred_img"], alone["pred_img"]),
                 "psnr_second_iteration_same_inputs_db": psnr(forced[0][0], forced[1][0]),
                 "psnr_second_iteration_disparity_same_inputs_db": psnr(forced[0][1], forced[1][1]),
                 "note": "sharded vs un-sharded rollout on the same rank, same device draws; the second iteration "
                         "free-runs from bf16-different first-iteration frames"}
    if rank != 0:
        return
    line = {"metric": "rollout_frames_per_second", "n_gpus": world, "sharded_vs_alone": check, "value": best["frames"] / (best["total_ms"] / 1e3), "unit": "frames/s",
            "iterations": args.iters, "frames": best["frames"], "total_ms": best["total_ms"], "host_wall_ms": best["wall_ms"],
            "generate_i2v_ms": best["i2v_ms"], "feedback_ms": best["feedback_ms"], "gpu_launches": best["launches"],
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"generate(): {args.iters} iteration(s) of 8 units at 384x512, {args.layers} blocks, "
                                   f"3 stages x {args.steps} steps, VAE encode + 2 decodes per iteration, device feedback"}}
    if args.cpu_feedback:
        from oracle import rollout_ref
        fbk = rollout_ref.Feedback(None)
        im = res["pred_img"][:, :, :57].float().cpu()
        dp = res["pred_disparity"][:, :, :57].float().cpu() * 2 - 1
        t3, t2 = res["trans3d"][:, :8].cpu(), res["trans2d"][:, :8].cpu()
        t0 = time.perf_counter()
        _, d = fbk.absorb(0, im, dp, t3, t2, ["w"] * 8)
        last = im[0, :, -25:].permute(1, 2, 3, 0)
        (torch.clamp(last * 0.5 + 0.5, 0, 1).to(torch.float32) * 255).to(torch.uint8)
        dd = d[:, :, -25:]
        s = 1 / dd[:, :, 0].max()
        torch.sqrt(dd * s * 0.95) * 2 - 1
        rollout_ref.raymap_to_pose(torch.randn(1, 6, 7, 48, 64))
        rollout_ref.camera_raymap(t2[:, -4:], t3[:, -4:], (384, 512), 8)
        line["cpu_feedback_ms"] = (time.perf_counter() - t0) * 1e3
        line["cpu_feedback_note"] = ("oracle arithmetic of the same feedback step on the host cores (no PIL, no D2H/H2D, "
                                     "history encode excluded)")
    os.write(out_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
