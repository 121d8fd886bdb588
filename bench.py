#!/usr/bin/env python
"""bench.py — frames/s of the DeepVerse denoise + VAE-decode hot path on B200.

    python bench.py --gpus 1 --steps K --warmup W              (ours, sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W      (the reference algorithm on the host CPU)
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): ONE autoregressive unit at the demo
shape 384x512 — full-depth (24-block) MMDiT, CFG batch 2, unit-1 token layout, 3 pyramid stages
x 5 Euler steps = 15 denoiser forwards with fused CFG + scheduler steps and two stage
transitions — followed by the tiled VAE decode of the RGB and the disparity latents
[1,16,2,48,64] -> [1,3,9,384,512] each.  A step emits 8 generated frames (frame 0 is the input
image).  `--workload iteration` runs a steady-state iteration instead (4 units, CFG batch 3 with
a history frame, two 57-frame decodes, 32 emitted frames).  Random-init weights, synthetic
latents/embeddings (no checkpoints offline).

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the
same through the public pipeline API with host buffers (H2D of the step's inputs and D2H of the
decoded frames inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "generated frames/sec (denoise + VAE decode), 384x512, 5 steps x 3 stages"
UNIT = "frames/s"
STEPS_PER_STAGE = [5, 5, 5]
LAT_H, LAT_W = 48, 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="unit", choices=["unit", "iteration"])
    ap.add_argument("--layers", type=int, default=24, help="debug only; anything but 24 is not the benchmark")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--group", type=int, default=0,
                    help="GPUs that share ONE rollout (CFG branches + VAE tiles sharded over them); the world is "
                         "split into world/group independent rollouts.  0 = 2 when the world is even, else 1")
    ap.add_argument("--profile-dump", default=None, help="write the per-launch CSV of the profiled step here")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload description (shared by both arms)
# ------------------------------------------------------------------------------------------------
def unit_layouts(workload):
    """Per unit: (n_branch, has_history, [per-stage clip dims oldest-first, EXCLUDING the noisy clip])."""
    s0, s1, s2 = (12, 16), (24, 32), (48, 64)
    if workload == "unit":
        # unit 1 of the first iteration: one condition frame at the stage's size
        return [dict(n_branch=2, hist=False, conds=[[(1, *s0)], [(1, *s1)], [(1, *s2)]], lat_T=2)]
    units = []
    # steady iteration (pipeline.py:626-658 with firstframe_mask = 0): unit k sees n = 4 + k clean
    # frames; stage 0/1: older frames at stage-0 size + the last one at the stage's size; stage 2:
    # all but the last two at stage 0, frame -2 at stage 1, the last at stage 2 (App. B: 240/528/1824
    # ... 384/672/1968 video tokens), CFG batch 3 + history frame
    for k in range(4):
        n = 4 + k
        units.append(dict(n_branch=3, hist=True, conds=[
            [(n - 1, *s0), (1, *s0)],
            [(n - 1, *s0), (1, *s1)],
            [(n - 2, *s0), (1, *s1), (1, *s2)],
        ], lat_T=8))
    return units


def frames_per_step(workload):
    return 8 if workload == "unit" else 32


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        load = [s for s, p in zip(sm, pw) if p > 300] or sm
        load.sort()
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# synthetic inputs (host, pinned when CUDA is there)
# ------------------------------------------------------------------------------------------------
def make_inputs(workload, dtype, pin):
    import torch
    g = torch.Generator().manual_seed(666)  # run.py:379 default seed
    units = []
    for u in unit_layouts(workload):
        B = u["n_branch"]
        d = dict(u)
        d["latents"] = (torch.randn(1, 38, 1, 12, 16, generator=g) * 2).to(dtype)  # x2 per bilinear halving, pipeline.py:557
        d["cond_tensors"] = [[torch.randn(B, 38, t, h, w, generator=g).to(dtype) for (t, h, w) in st] for st in u["conds"]]
        d["block_noise"] = [torch.randn(1, 38, 1, 24, 32, generator=g).to(dtype),
                            torch.randn(1, 38, 1, 48, 64, generator=g).to(dtype)]
        d["enc"] = torch.randn(B, 77, 4096, generator=g).to(dtype)
        d["pooled"] = torch.randn(B, 2048, generator=g)
        mask = torch.zeros(B, 77, dtype=torch.long)
        mask[0, :1] = 1
        mask[1:, :12] = 1
        d["mask"] = mask
        d["history"] = torch.randn(1, 38, 1, 48, 64, generator=g).to(dtype) if u["hist"] else None
        units.append(d)
    lat_T = units[-1]["lat_T"]
    dec = [torch.randn(1, 16, lat_T, LAT_H, LAT_W, generator=g).to(dtype) for _ in range(2)]  # rgb, disparity
    if pin:
        def P(t):
            return t.pin_memory() if t is not None else None
        for d in units:
            d["latents"] = P(d["latents"])
            d["cond_tensors"] = [[P(c) for c in st] for st in d["cond_tensors"]]
            d["block_noise"] = [P(c) for c in d["block_noise"]]
            d["enc"], d["pooled"], d["mask"], d["history"] = P(d["enc"]), P(d["pooled"]), P(d["mask"]), P(d["history"])
        dec = [P(t) for t in dec]
    return units, dec


def nbytes(x):
    import torch
    if x is None:
        return 0
    if isinstance(x, torch.Tensor):
        return x.numel() * x.element_size()
    if isinstance(x, (list, tuple)):
        return sum(nbytes(i) for i in x)
    return 0


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, layers):
    """A bounded sample of the same workload on the CPU, FLOP-scaled to a whole step:
    one MMDiT forward per pyramid stage (of 5 each) at the step's first layout, and the decode of one
    16x16-latent window pair of the VAE (1/4 of a tile's area), scaled by algorithmic FLOPs."""
    import torch
    from oracle import mmdit_ref, vae_ref, weights
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    units, dec = make_inputs(workload, torch.float32, pin=False)
    cfg, W = weights.mmdit_weights(dict(num_layers=layers), seed=1)
    pos = mmdit_ref.sincos_2d_table(1536, 192, 64)
    u = units[0]
    t_forward = []
    with torch.no_grad():
        for i_s in range(3):
            h, w = 12 * 2 ** i_s, 16 * 2 ** i_s
            x = torch.cat([torch.randn(1, 38, 1, h, w)] * u["n_branch"])
            hist = torch.cat([u["history"]] * 3) if u["history"] is not None else None
            hmask = torch.cat([torch.zeros(2, 192), torch.ones(1, 192)]) if hist is not None else None
            t0 = time.perf_counter()
            mmdit_ref.mmdit_forward(W, cfg, list(u["cond_tensors"][i_s]) + [x], torch.full((u["n_branch"],), 500.0),
                                    u["enc"], u["mask"], u["pooled"], hist, hmask, 2 if hist is not None else None,
                                    pos_table=pos)
            t_forward.append(time.perf_counter() - t0)
    del W
    vcfg, VW = weights.vae_weights(None, seed=2)
    z = torch.randn(1, 16, 2, 16, 16)
    with torch.no_grad():
        t0 = time.perf_counter()
        vae_ref.chunk_decode(VW, vcfg, z, 1)
        t_vae = time.perf_counter() - t0
    # FLOP scaling of the VAE sample to two full tiled decodes of lat_T frames
    lat_T = units[-1]["lat_T"]
    per_area = t_vae / (16 * 16)                    # first window (2 latent frames -> 9 frames)
    tiles_area = sum(th * tw for th in (32, 24) for tw in (32, 32, 16))
    win_scale = 1.0 + (lat_T - 2) * (8.0 / 9.0)     # later windows emit 8 frames each
    t_decode = 2 * per_area * tiles_area * win_scale
    n_units = len(units)
    t_unit = sum(5 * t for t in t_forward)
    step_s = n_units * t_unit + t_decode
    sample = (f"oracle port: 1 of 5 MMDiT forwards per stage ({', '.join(f'{t:.2f}s' for t in t_forward)}) "
              f"+ VAE chunk_decode of a [1,16,2,16,16] latent ({t_vae:.2f}s), FLOP/area-scaled to "
              f"{n_units} unit(s) + 2 tiled decodes of {lat_T} latent frames")
    return step_s, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = []
    sample = ""
    cores = 0
    for i in range(max(1, min(args.steps, 2)) + min(args.warmup, 1)):
        step_s, cores, sample = cpu_reference_sample(args.workload, args.layers)
        t_all.append(step_s)
    t_all = t_all[min(args.warmup, 1):]
    step_s = sum(t_all) / len(t_all)
    val = frames_per_step(args.workload) / step_s
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic", "config": workload_config(args, None),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, extra):
    name = ("C2: one AR unit (unit-1 layout, CFG batch 2, 3 stages x 5 steps = 15 full-depth MMDiT forwards) + "
            "tiled VAE decode of RGB and disparity latents [1,16,2,48,64] -> 2 x [1,3,9,384,512]; 8 frames/step"
            if args.workload == "unit" else
            "C4 steady iteration: 4 AR units (CFG batch 3 + history frame, 60 forwards) + tiled VAE decode of RGB "
            "and disparity latents [1,16,8,48,64] -> 2 x [1,3,57,384,512]; 32 emitted frames/step")
    cfg = {"workload": name, "resolution": "384x512", "mmdit_layers": args.layers, "weights": "random-init (seeded)",
           "l2": "per-step working set (4.1 GB bf16 weights + activations) exceeds the 126 MB L2"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepv_b200 import _lib
    from deepv_b200.mmdit import B200MMDiT
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.scheduler import B200Scheduler
    from deepv_b200.vae import B200VAE
    from deepv_b200 import synthetic as synth  # seeded random-init weights (no checkpoints offline)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sm_100a path has no CPU fallback "
                         "(use --impl reference for the host-CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    dtype = torch.bfloat16

    cfg, W = synth.mmdit_weights(dict(num_layers=args.layers), seed=1)
    dit = B200MMDiT(W, cfg, device=dev)
    del W
    vcfg, VW = synth.vae_weights(None, seed=2)
    vae = B200VAE(VW, vcfg, device=dev, dtype=dtype)
    vae.enable_tiling()
    del VW
    pipe = B200Pipeline(dit, vae, B200Scheduler(num_train_timesteps=1000, shift=1.0, stages=3,
                                                stage_range=[0, 1 / 3, 2 / 3, 1], gamma=0.3333),
                        device=dev, torch_dtype=dtype)
    units_h, dec_h = make_inputs(args.workload, dtype, pin=True)

    def to_dev(units, dec):
        ud = []
        for d in units:
            e = dict(d)
            e["latents"] = d["latents"].to(dev, non_blocking=True)
            e["cond_tensors"] = [[c.to(dev, non_blocking=True) for c in st] for st in d["cond_tensors"]]
            e["block_noise"] = [c.to(dev, non_blocking=True) for c in d["block_noise"]]
            e["enc"] = d["enc"].to(dev, non_blocking=True)
            e["pooled"] = d["pooled"].to(dev, non_blocking=True)
            e["mask"] = d["mask"].to(dev, non_blocking=True)
            e["history"] = d["history"].to(dev, non_blocking=True) if d["history"] is not None else None
            ud.append(e)
        return ud, [t.to(dev, non_blocking=True) for t in dec]

    from deepv_b200.parallel import Shard
    gsz = args.group if args.group > 0 else (2 if world % 2 == 0 else 1)
    shard = Shard.grouped(gsz) if world > 1 else Shard(0, 1, None)
    n_rollouts = world // gsz if world > 1 else 1
    if world > 1:
        shard.setup_sp(dev)   # CFG branch groups x Ulysses ranks inside every rollout group

    def step(units, dec, fetch):
        outs = []
        for d in units:
            # one rollout sharded over the ranks: CFG branches on different GPUs when there are enough
            # ranks (all-gather of the branch predictions per step), else the whole CFG batch locally
            sh = shard if shard.active else None
            lat = pipe.generate_one_unit(d["latents"], d["history"], d["cond_tensors"], d["enc"], d["mask"],
                                         d["pooled"], STEPS_PER_STAGE, block_noise=d["block_noise"], shard=sh)
            outs.append(lat[-1])
        # the rollout decodes the first 16 (RGB) and next 16 (disparity) latent channels of all units
        # (pipeline.py:686-696); synthetic latents of that shape keep the decode at [1,16,lat_T,48,64]
        if shard.active:
            img, dsp = pipe.decode_latents_sharded([dec[0], dec[1]], shard)  # tiles x modalities over ranks
        else:
            img = pipe.decode_latent(dec[0])
            dsp = pipe.decode_latent(dec[1])
        if fetch:
            return img.to("cpu", non_blocking=True), dsp.to("cpu", non_blocking=True)
        return img, dsp

    units_d, dec_d = to_dev(units_h, dec_h)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing (value) -----------------------------------------------------------
    for _ in range(args.warmup):
        step(units_d, dec_d, False)
    barrier()
    lib.dv_launch_count_reset()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(units_d, dec_d, False)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = lib.dv_launch_count() // args.steps
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    # one rollout per group of `gsz` GPUs (sharded inside the group), world/gsz rollouts in all
    frames = frames_per_step(args.workload) * n_rollouts
    value = frames / (ms / 1e3)

    # ---- end-to-end through the public API with host buffers -------------------------------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(nbytes(d[k]) for d in units_h for k in ("latents", "cond_tensors", "block_noise", "enc", "pooled", "mask", "history")) + nbytes(dec_h)
        for _ in range(2):
            u, dd = to_dev(units_h, dec_h)
            r = step(u, dd, True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            u, dd = to_dev(units_h, dec_h)
            r = step(u, dd, True)
            torch.cuda.current_stream().synchronize()  # the frames are on the host before the next step
        e1.record()
        barrier()
        ms_e = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms_e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = t.item()
        e2e = {"value": frames / (ms_e / 1e3), "unit": UNIT, "ms_per_step": ms_e, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": nbytes(r[0]) + nbytes(r[1])}

    # ---- roofline of the dominant kernel (tcgen05 GEMM / implicit-GEMM conv) ---------------------
    lib.dv_profile_enable(1)
    lib.dv_profile_reset()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    step(units_d, dec_d, False)
    p1.record()
    torch.cuda.synchronize()
    lib.dv_profile_enable(0)
    prof_ms = p0.elapsed_time(p1)
    kinds = {}
    for k, name in ((0, "gemm_dense"), (1, "gemm_conv"), (2, "attention")):
        cnt, kms, fl, by = C.c_longlong(), C.c_double(), C.c_double(), C.c_double()
        _lib.check(lib.dv_profile_summary(k, C.byref(cnt), C.byref(kms), C.byref(fl), C.byref(by)))
        kinds[name] = dict(launches=cnt.value, ms=kms.value, tflop=fl.value / 1e12,
                           tflops=(fl.value / 1e12) / (kms.value / 1e3) if kms.value > 0 else 0.0)
    if args.profile_dump and rank == 0:
        _lib.check(lib.dv_profile_dump(args.profile_dump.encode()), "dv_profile_dump")
    lib.dv_profile_reset()
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    gk_ms = kinds["gemm_dense"]["ms"] + kinds["gemm_conv"]["ms"]
    gk_tf = kinds["gemm_dense"]["tflop"] + kinds["gemm_conv"]["tflop"]
    gk_n = kinds["gemm_dense"]["launches"] + kinds["gemm_conv"]["launches"]
    achieved = gk_tf / (gk_ms / 1e3) if gk_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel / gemm_pair_kernel (dense) + conv_halo_kernel / implicit-GEMM conv3d (tcgen05/TMEM/TMA)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "frac_of_nominal_2250": achieved / 2250.0, "peak_source": peak_src, "traffic": None,
                # `achieved` aggregates ~2200 launches of many shapes, so one per-launch DRAM figure does not
                # exist; the largest single shape as captured by `ncu --set full` (profiles/r01_ncu_conv_halo_summary.txt):
                "traffic_sample": {"launch": "conv_halo_kernel 128->128 3x3x3 @ [9][256][256] (20 launches/step)",
                                   "dram_bytes": 285.0e6, "algorithmic_bytes": 302.9e6,
                                   "tensor_pipe_active_pct": 84.4,
                                   "source": "profiles/r01_ncu_conv_halo_summary.txt"},
                "launches_per_step": gk_n, "avg_launch_us": gk_ms * 1e3 / gk_n if gk_n else None,
                "kernel_share_of_step": gk_ms / prof_ms if prof_ms > 0 else None, "by_kind": kinds,
                "how": "CUDA events on the launching stream around every launch of one extra profiled step"}

    # ---- CPU baseline (rank 0, N = 1) --------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step_s, cores, sample = cpu_reference_sample(args.workload, args.layers)
        cpu = {"value": frames_per_step(args.workload) / step_s, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": sample}

    if rank == 0:
        bad = [r for r in clk.get("reasons", []) if r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, {"parallelism": (
                    f"{n_rollouts} rollout(s) x {gsz} GPU(s) each: CFG branches on sub-groups (all-gather of the branch "
                    f"predictions per step) x Ulysses sequence parallelism inside a branch (all-to-all of heads<->tokens "
                    f"around every attention, NCCL), VAE tiles x modalities dealt over the group (tile broadcast)"
                    if world > 1 else "single GPU")}),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "clock_rejected": bool(bad),
                "roofline": roofline, "cpu_baseline": cpu,
                "published_reference": {"value": 4.0, "unit": "frames/s", "hardware": "1x A800, full run.py pipeline",
                                        "source": "README.md:78", "comparable": False}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # libraries (NCCL banners, ...) may write to fd 1: keep the real stdout for the ONE JSON line
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
