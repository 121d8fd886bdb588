#!/usr/bin/env python
"""bench.py — frames/s of the DeepVerse denoise + VAE hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W              (ours, sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W      (the reference's own code on the host CPU)
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, ONE rollout over all of them)

Workloads (random-init seeded weights, synthetic image / embeddings: no checkpoints offline):
  rollout (default)  one `generate()` of the reference's run.py configuration (pipeline.py:264-424) at 384x512:
                     iteration 0 = C3 (one input frame, 8 autoregressive units, CFG batch 2, 120 full-depth MMDiT
                     forwards, 57 frames) + iteration 1 = C4 steady state (25 input frames + disparity + ray map +
                     history frame, 4 units, CFG batch 3, 60 forwards, 32 new frames), with the VAE encodes, both
                     57-frame tiled decodes per iteration and the device feedback between them: 89 emitted frames.
  unit               C2: one autoregressive unit (15 forwards, unit-1 layout) + two 9-frame decodes (round-1 headline).

One JSON line on stdout (rank 0).  `value` = emitted frames/s with the inputs resident in HBM; `e2e` = the same
through the public API with HOST buffers (H2D of the step's inputs, D2H of the frames inside the timed region).
`roofline` = the dominant kernel (largest share of the step) + every kernel class against its own bound;
`cpu_baseline` / `--impl reference` = the UNMODIFIED reference files (staged under baseline/_ref, imported through
oracle/_shim.py) on the host cores, a bounded sample FLOP-scaled to the workload; `same_box_eager` = the same
reference modules in bf16 eager on this GPU (informational).
"""
from __future__ import annotations

import argparse
import csv
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "generated frames/sec (denoise + VAE decode), 384x512, 5 steps x 3 stages"
UNIT = "frames/s"
STEPS_PER_STAGE = [5, 5, 5]
LAT_H, LAT_W = 48, 64
SCHED_KW = dict(num_train_timesteps=1000, shift=1.0, stages=3, stage_range=[0, 1 / 3, 2 / 3, 1], gamma=0.3333)  # run.py:27-31
ROLLOUT_ITERS = 2
ROLLOUT_PROMPTS = (["w", "a", "d", "w"] * 3)[:8 + 4 * (ROLLOUT_ITERS - 1)]      # 12 actions -> 2 iterations (pipeline.py:276-279)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rollout", choices=["rollout", "unit"])
    ap.add_argument("--layers", type=int, default=24, help="debug only; anything but 24 is not the benchmark")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-same-box-eager", action="store_true")
    ap.add_argument("--no-replicas", action="store_true", help="N > 1: skip the extra pass that times N independent rollouts")
    ap.add_argument("--group", type=int, default=0,
                    help="GPUs that share ONE rollout (CFG branches x Ulysses ranks + VAE tiles sharded over them); the "
                         "world is split into world/group rollouts.  0 = the whole world (one rollout over all GPUs)")
    ap.add_argument("--profile-dump", default=None, help="keep the per-launch CSV of the profiled step here")
    return ap.parse_args()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else any library or the reference's own
    modules print (NCCL banners, `print` in mmdit.py:1249) was redirected to stderr at start-up."""
    out = _REAL_STDOUT or sys.__stdout__
    out.write(json.dumps(line) + "\n")
    out.flush()


def frames_per_step(workload):
    return 8 if workload == "unit" else 57 + 32 * (ROLLOUT_ITERS - 1)


def workload_name(workload):
    if workload == "unit":
        return ("C2: one AR unit (unit-1 layout, CFG batch 2, 3 stages x 5 steps = 15 full-depth MMDiT forwards) + "
                "tiled VAE decode of RGB and disparity latents [1,16,2,48,64] -> 2 x [1,3,9,384,512]; 8 frames/step")
    return (f"generate() rollout, {len(ROLLOUT_PROMPTS)} action prompts = C3 first iteration (8 units, CFG batch 2, 120 forwards, "
            "57 frames) + C4 steady iteration (25 input frames, 4 units, CFG batch 3 + history frame, 60 forwards, 32 new "
            "frames); per iteration VAE encode of the inputs + history frame, two tiled decodes [1,16,8,48,64] -> "
            f"[1,3,57,384,512], device feedback; {frames_per_step('rollout')} emitted frames/step")


def workload_config(args, parallelism):
    return {"workload": workload_name(args.workload), "resolution": "384x512", "mmdit_layers": args.layers,
            "weights": "random-init (seeded)", "parallelism": parallelism,
            "l2": "per-step working set (4.1 GB bf16 MMDiT weights + 0.7 GB VAE weights + activations) exceeds the 126 MB L2"}


def parallelism_string(world, gsz):
    """What actually runs (deepv_b200.parallel.sp_ranks): branch groups x Ulysses ranks per CFG batch size."""
    if world == 1:
        return "single GPU"
    from deepv_b200.parallel import sp_ranks
    nb2, sp2 = sp_ranks(gsz, 2)
    nb3, sp3 = sp_ranks(gsz, 3)
    return (f"{world // gsz} rollout(s) x {gsz} GPU(s) each; inside a rollout: CFG batch 2 -> {nb2} branch group(s) x Ulysses "
            f"sequence parallelism {sp2} (heads<->tokens exchange around every attention{' - not active' if sp2 == 1 else ''}); "
            f"CFG batch 3 -> {nb3} branch group(s) x Ulysses {sp3}{' - not active' if sp3 == 1 else ''}; branch predictions "
            f"all-gathered per step when branch groups > 1; VAE tiles x modalities dealt over the {gsz} rank(s), tiles broadcast")


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
# synthetic inputs
# ------------------------------------------------------------------------------------------------
def prompt_table(dtype):
    """The file `model_cfg['text_embeds_path']` points to (pipeline.py:199), synthetic: SURVEY.md §8d."""
    import torch
    g = torch.Generator().manual_seed(3)
    return {k: dict(prompt_embeds=torch.randn(1, 77, 4096, generator=g).to(dtype),
                    pooled_prompt_embeds=torch.randn(1, 2048, generator=g),
                    prompt_attention_mask=(torch.arange(77) < n).long().view(1, 77))
            for k, n in (("empty", 1), ("w", 12), ("a", 12), ("d", 12))}


def first_frame():
    import torch
    g = torch.Generator().manual_seed(0)
    return (torch.rand(384, 512, 3, generator=g) * 255).to(torch.uint8)


def unit_inputs(dtype, pin):
    """C2 inputs on the host (pinned when CUDA is there)."""
    import torch
    g = torch.Generator().manual_seed(666)  # run.py:379 default seed
    s = [(12, 16), (24, 32), (48, 64)]
    d = dict(latents=(torch.randn(1, 38, 1, 12, 16, generator=g) * 2).to(dtype),   # x2 per bilinear halving, pipeline.py:557
             cond_tensors=[[torch.randn(2, 38, 1, h, w, generator=g).to(dtype)] for (h, w) in s],
             block_noise=[torch.randn(1, 38, 1, 24, 32, generator=g).to(dtype), torch.randn(1, 38, 1, 48, 64, generator=g).to(dtype)],
             enc=torch.randn(2, 77, 4096, generator=g).to(dtype), pooled=torch.randn(2, 2048, generator=g))
    mask = torch.zeros(2, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1:, :12] = 1
    d["mask"] = mask
    dec = [torch.randn(1, 16, 2, LAT_H, LAT_W, generator=g).to(dtype) for _ in range(2)]  # rgb, disparity
    if pin:
        def P(t):
            return t.pin_memory()
        d["latents"] = P(d["latents"])
        d["cond_tensors"] = [[P(c) for c in st] for st in d["cond_tensors"]]
        d["block_noise"] = [P(c) for c in d["block_noise"]]
        d["enc"], d["pooled"], d["mask"] = P(d["enc"]), P(d["pooled"]), P(d["mask"])
        dec = [P(t) for t in dec]
    return d, dec


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline / same-box eager: the reference's own modules on a bounded sample
# ------------------------------------------------------------------------------------------------
SAMPLE_CLIPS = {0: [(1, 12, 16), (1, 12, 16)], 1: [(1, 24, 32), (1, 24, 32)], 2: [(1, 48, 64), (1, 48, 64)]}   # unit-1 layouts
# CPU sample (a few seconds per step on 16 cores) / GPU sample (large enough not to be launch-bound in eager mode)
SAMPLE_SHAPES = {"cpu": dict(B=(2, 2, 1), z=(2, 8, 8), x=(1, 128, 128)),
                 "cuda": dict(B=(2, 2, 2), z=(8, 32, 32), x=(25, 256, 256))}


class ReferenceSampler:
    """One MMDiT forward per pyramid stage (of the 5 per stage) at the unit-1 layout, the VAE chunked decode of one
    latent tile and the VAE encode of one pixel tile (SAMPLE_SHAPES), through the reference's own MMDiT.forward
    (mmdit.py:1467-1530) and CausalVideoVAE.decode / .encode (vae.py:885-920,844-863).
    `step()` times that sample once; `scale()` FLOP-scales it to the workload (deepv_b200/work.py)."""

    def __init__(self, layers, device="cpu", dtype=None):
        import torch
        from deepv_b200 import synthetic as synth
        self.torch = torch
        self.device = torch.device(device)
        self.dtype = dtype or torch.float32
        self.layers = layers
        self.kind = "reference"
        cfg, W = synth.mmdit_weights(dict(num_layers=layers), seed=1)
        vcfg, VW = synth.vae_weights(None, seed=2, encoder=True)
        try:
            from oracle import reference_loader as rl
            if not rl.available():
                raise FileNotFoundError("reference sources not staged under baseline/_ref")
            self.dit = rl.build_mmdit(cfg, W, self.device, self.dtype)
            self.vae = rl.build_vae(vcfg, VW, self.device, self.dtype)
            self.note = "unmodified reference model/mmdit.py + model/vae.py (baseline/_ref through oracle/_shim.py)"
        except Exception as e:  # the restatement is the fallback checker-grade port
            if self.device.type != "cpu":
                raise
            from oracle import mmdit_ref, vae_ref
            self.kind = "port"
            self.note = f"oracle port (reference files unavailable: {e})"
            self.dit = self.vae = None
            self._port = (mmdit_ref, vae_ref, cfg, W, vcfg, VW, mmdit_ref.sincos_2d_table(1536, 192, 64))
        g = torch.Generator().manual_seed(667)
        mk = lambda *s: torch.randn(*s, generator=g).to(self.device, self.dtype)   # noqa: E731
        self.shapes = SAMPLE_SHAPES[self.device.type]
        self.clips = {st: [mk(self.shapes["B"][st], 38, t, h, w) for (t, h, w) in cl] for st, cl in SAMPLE_CLIPS.items()}
        self.enc, self.pooled = mk(2, 77, 4096), mk(2, 2048)
        self.mask = torch.zeros(2, 77, dtype=torch.long, device=self.device)
        self.mask[0, :1] = 1
        self.mask[1, :12] = 1
        self.z = mk(1, 16, *self.shapes["z"])
        self.x = mk(1, 3, *self.shapes["x"])

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize()

    def _forward(self, st):
        torch = self.torch
        B = self.shapes["B"][st]
        t = torch.full((B,), 500.0, device=self.device, dtype=self.dtype)
        enc, mask, pooled = self.enc[2 - B:], self.mask[2 - B:], self.pooled[2 - B:]    # B = 1: the prompted branch
        if self.dit is None:
            mr, _, cfg, W, _, _, pos = self._port
            return mr.mmdit_forward(W, cfg, self.clips[st], t.float(), enc, mask, pooled, pos_table=pos)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.device.type == "cuda"):   # pipeline.py:486
            return self.dit(sample=[list(self.clips[st])], timestep_ratio=t, encoder_hidden_states=enc,
                            encoder_attention_mask=mask, pooled_projections=pooled)[0]

    def _decode(self):
        if self.vae is None:
            _, vr, _, _, vcfg, VW, _ = self._port
            return vr.chunk_decode(VW, vcfg, self.z, 1)
        return self.vae.decode(self.z, temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample   # pipeline.py:713

    def _encode(self):
        if self.vae is None:
            _, vr, _, _, vcfg, VW, _ = self._port
            return vr.tiled_encode(VW, vcfg, self.x)
        return self.vae.encode(self.x).latent_dist.mean

    def step(self):
        """Seconds of each part of the sample."""
        torch = self.torch
        out = {}
        with torch.no_grad():
            for name, fn in (("fwd0", lambda: self._forward(0)), ("fwd1", lambda: self._forward(1)),
                             ("fwd2", lambda: self._forward(2)), ("dec", self._decode), ("enc", self._encode)):
                self._sync()
                t0 = time.perf_counter()
                fn()
                self._sync()
                out[name] = time.perf_counter() - t0
        return out

    def scale(self, t, workload):
        """FLOP-scale one sample's seconds to a whole step of `workload`."""
        from deepv_b200 import work
        f_s = [work.mmdit_flops(self.shapes["B"][st], SAMPLE_CLIPS[st], False, n_layers=self.layers)["total"] for st in range(3)]
        f_dec, f_enc = work.vae_decode_tile_flops(*self.shapes["z"]), work.vae_encode_flops(*self.shapes["x"])
        if workload == "unit":
            stage = [5 * f_s[0], 5 * f_s[1], 5 * f_s[2]]
            dec, enc = 2 * work.vae_decode_flops(2), 0.0
        else:
            stage = [0.0, 0.0, 0.0]
            for f in work.rollout_forwards(ROLLOUT_ITERS, LAT_H, LAT_W, STEPS_PER_STAGE):
                stage[f["stage"]] += f["count"] * work.mmdit_flops(f["B"], f["clips"], f["hist"], n_layers=self.layers)["total"]
            rw = work.rollout_work(ROLLOUT_ITERS, LAT_H, LAT_W, STEPS_PER_STAGE)
            dec, enc = rw["vae_decode"], rw["vae_encode"]
        sec = sum(stage[i] / f_s[i] * t[f"fwd{i}"] for i in range(3)) + dec / f_dec * t["dec"] + enc / f_enc * t["enc"]
        return sec

    def describe(self, t):
        sh = self.shapes
        return (f"{self.note}: 1 MMDiT forward per stage at the unit-1 layout, CFG batch {sh['B']} ({t['fwd0']:.3f}s, "
                f"{t['fwd1']:.3f}s, {t['fwd2']:.3f}s) + VAE chunked decode of a [1,16,{sh['z'][0]},{sh['z'][1]},{sh['z'][2]}] latent "
                f"({t['dec']:.3f}s) + VAE encode of a [1,3,{sh['x'][0]},{sh['x'][1]},{sh['x'][2]}] clip ({t['enc']:.3f}s); "
                f"FLOP-scaled per stage / per decode / per encode to the workload (deepv_b200/work.py)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    rs = ReferenceSampler(args.layers, "cpu", torch.float32)
    times = []
    for i in range(args.warmup + args.steps):
        t = rs.step()
        if i >= args.warmup:
            times.append(t)
    mean = {k: sum(t[k] for t in times) / len(times) for k in times[0]}
    step_s = rs.scale(mean, args.workload)
    val = frames_per_step(args.workload) / step_s
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "sample_ms_per_step": sum(mean.values()) * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "fp32",
            # the same `config` as our arm at this N (the arm itself runs on the host cores: `cpu_baseline.cores`)
            "data": "synthetic", "config": workload_config(args, parallelism_string(world, args.group if args.group > 0 else world)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": rs.kind, "sample": rs.describe(mean)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def same_box_eager(args):
    """The reference's own modules in bf16 eager on this GPU (pipeline.py:190-191,486: bf16 weights, autocast), same
    bounded sample, same FLOP scaling.  Informational: the reference's GPU path, not its CPU path."""
    import torch
    try:
        rs = ReferenceSampler(args.layers, "cuda", torch.bfloat16)
        for _ in range(2):
            rs.step()
        reps = [rs.step() for _ in range(3)]
        best = {k: min(r[k] for r in reps) for k in reps[0]}
        sec = rs.scale(best, args.workload)
        out = {"value": frames_per_step(args.workload) / sec, "unit": UNIT, "ms_per_step": sec * 1e3,
               "sample": rs.describe(best), "dtype": "bf16 (autocast, as pipeline.py:486)",
               "note": "torch %s eager: cuBLAS / cuDNN / SDPA as the reference reaches them; excludes the reference's CPU "
                       "block-noise loop (pipeline.py:431-437, ~1 s per unit) and its host round trips" % torch.__version__}
        del rs
        torch.cuda.empty_cache()
        return out
    except Exception as e:  # never fail the bench because of the informational bar
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------
# per-launch profile -> roofline classes
# ------------------------------------------------------------------------------------------------
def classify(kind, tag, n_sm=148):
    if kind == 1:
        return "conv"
    if kind == 2:
        return "attention"
    if kind == 0:
        if tag.startswith("pbk"):      # persistent block kernel: only used for layouts with fewer tiles than SMs
            return "dense_small_M"
        m = re.match(r"gemm B(\d+) M(\d+)(?:\+(\d+))? N(\d+) K(\d+)", tag)
        if m:
            B, m0, m1, N = int(m.group(1)), int(m.group(2)), int(m.group(3) or 0), int(m.group(4))
            tiles = B * (-(-m0 // 128) + (-(-m1 // 128) if m1 else 0)) * -(-N // 256)
            return "dense_large_M" if tiles >= n_sm else "dense_small_M"
        return "dense_large_M"
    return "hbm_elementwise"


def roofline_from_profile(path, prof_ms, peaks):
    rows = list(csv.DictReader(open(path)))
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_bw = peaks.get("hbm_gbs") or 6500.0
    src = ("MEASURED_PEAKS.json (bf16_tflops_sustained: kernels timed inside a long step; hbm_gbs: measured copy bandwidth)"
           if peaks else "fallback of B200_PROFILING.md: 1.4 PFLOP/s sustained bf16, 6.5 TB/s HBM")
    classes, tags = {}, {}
    for r in rows:
        ms = float(r["ms"])
        if ms < 0:
            continue
        c = classify(int(r["kind"]), r["tag"])
        a = classes.setdefault(c, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        t = tags.setdefault((c, r["tag"]), dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        for d in (a, t):
            d["launches"] += 1
            d["ms"] += ms
            d["flops"] += float(r["flops"])
            d["bytes"] += float(r["bytes"])
    total_ms = sum(a["ms"] for a in classes.values())
    out = {}
    for c, a in classes.items():
        tf = a["flops"] / 1e12 / (a["ms"] / 1e3) if a["ms"] > 0 else 0.0
        gb = a["bytes"] / 1e9 / (a["ms"] / 1e3) if a["ms"] > 0 else 0.0
        bound = "hbm" if c in ("dense_small_M", "hbm_elementwise") else "tensor"
        e = {"bound": bound, "launches": a["launches"], "ms": round(a["ms"], 3), "share_of_kernel_time": round(a["ms"] / total_ms, 4),
             "avg_launch_us": round(a["ms"] * 1e3 / a["launches"], 2)}
        if bound == "tensor":
            e.update(achieved=round(tf, 1), peak=peak_tf, unit="TFLOP/s", frac=round(tf / peak_tf, 4),
                     frac_of_nominal_2250=round(tf / 2250.0, 4))
        else:
            e.update(achieved=round(gb, 1), peak=peak_bw, unit="GB/s", frac=round(gb / peak_bw, 4))
            if c == "dense_small_M":
                e.update(tflops=round(tf, 1), note="fewer 128x256 tiles than SMs: bounded by streaming the weights once (bytes = "
                                                   "A + W + out of every launch), not by the tensor pipe")
        out[c] = e
    dom_class = max(classes.items(), key=lambda kv: kv[1]["ms"])[0]
    (dc, dtag), d = max(((k, v) for k, v in tags.items() if k[0] == dom_class), key=lambda kv: kv[1]["ms"])
    per_launch_ms = d["ms"] / d["launches"]
    bound = out[dc]["bound"]
    traffic, traffic_src = None, None
    try:
        tr = json.loads((ROOT / "profiles" / "r02_ncu_traffic.json").read_text())
        hit = tr.get(dtag)
        if hit:
            traffic, traffic_src = hit["dram_bytes_per_launch"], hit.get("source")
    except Exception:
        pass
    if bound == "tensor":
        ach, peak, unit = d["flops"] / d["launches"] / 1e12 / (per_launch_ms / 1e3), peak_tf, "TFLOP/s"
    else:
        ach, peak, unit = d["bytes"] / d["launches"] / 1e9 / (per_launch_ms / 1e3), peak_bw, "GB/s"
    top = {"bound": bound, "kernel": dtag, "kernel_class": dc, "achieved": round(ach, 1), "peak": peak, "unit": unit,
           "frac": round(ach / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
           "frac_of_burst_peak": round(ach / peaks["bf16_tflops"], 4) if bound == "tensor" and peaks.get("bf16_tflops") else None,
           "frac_of_nominal_2250": round(ach / 2250.0, 4) if bound == "tensor" else None,
           "algorithmic_flops_per_launch": d["flops"] / d["launches"], "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
           "avg_launch_us": round(per_launch_ms * 1e3, 2), "launches_per_step": d["launches"],
           "kernel_share_of_step": round(d["ms"] / prof_ms, 4) if prof_ms > 0 else None,
           "peak_source": src, "classes": out, "kernel_time_ms": round(total_ms, 2), "profiled_step_ms": round(prof_ms, 2),
           "executed_tflop": round(sum(a["flops"] for a in classes.values()) / 1e12, 1),
           "how": "CUDA events on the launching stream around every launch of one extra profiled step (graph replay and "
                  "programmatic-dependent-launch overlap are off while profiling); dominant kernel = the launch shape with the "
                  "largest total time inside the kernel class with the largest share of the step"}
    return top


def whole_step_summary(roofline, ms, gsz, peaks, rollout, layers):
    """The whole step against the tensor roofline: algorithmic FLOPs of the workload (deepv_b200/work.py)
    over the timed step of this rank's rollout group, next to the FLOPs the profiled step actually executed."""
    from deepv_b200 import work
    if rollout:
        rw = work.rollout_work(ROLLOUT_ITERS, LAT_H, LAT_W, STEPS_PER_STAGE)
        alg, exe_model = rw["total"], rw["executed"]
    else:
        alg = 5 * sum(work.mmdit_flops(2, SAMPLE_CLIPS[st], False, n_layers=layers)["total"] for st in range(3)) \
            + 2 * work.vae_decode_flops(2)
        exe_model = alg
    sus = peaks.get("bf16_tflops_sustained") or 1400.0
    return {"algorithmic_tflop": round(alg / 1e12, 1), "tflops_per_gpu": round(alg / 1e12 / (ms / 1e3) / gsz, 1),
            "frac_of_sustained_peak": round(alg / 1e12 / (ms / 1e3) / gsz / sus, 4), "gpus_per_rollout": gsz,
            "note": "algorithmic_tflop is the REFERENCE's work (deepv_b200/work.py); a continuation iteration decodes 25 frames per "
                    "video that the reference throws away and the trimmed decode never computes, so the FLOPs this rank actually "
                    "executed are roofline.executed_tflop",
            "executed_tflop_work_model": round(exe_model / 1e12, 1),
            "executed_frac_of_sustained_peak": round(roofline.get("executed_tflop", 0.0) / (ms / 1e3) / sus, 4)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepv_b200 import _lib
    from deepv_b200 import synthetic as synth  # seeded random-init weights (no checkpoints offline)
    from deepv_b200.mmdit import B200MMDiT
    from deepv_b200.parallel import Shard
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.rollout import B200Rollout, DeviceNoise, PromptCache
    from deepv_b200.scheduler import B200Scheduler
    from deepv_b200.vae import B200VAE

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
    rollout = args.workload == "rollout"

    cfg, W = synth.mmdit_weights(dict(num_layers=args.layers), seed=1)
    dit = B200MMDiT(W, cfg, device=dev)
    del W
    vcfg, VW = synth.vae_weights(None, seed=2, encoder=rollout)
    vae = B200VAE(VW, vcfg, device=dev, dtype=dtype)
    vae.enable_tiling()
    del VW
    pipe = B200Pipeline(dit, vae, B200Scheduler(**SCHED_KW), model_cfg=dict(num_inference_steps=STEPS_PER_STAGE[0]),
                        device=dev, torch_dtype=dtype)

    gsz = args.group if args.group > 0 else world
    shard = Shard.grouped(gsz) if world > 1 else Shard(0, 1, None)
    n_rollouts = world // gsz if world > 1 else 1
    if world > 1:
        shard.setup_sp(dev)   # CFG branch groups x Ulysses ranks inside every rollout group
    sh = shard if shard.active else None

    if rollout:
        ro = B200Rollout(pipe, PromptCache(prompt_table(dtype), None, dev))
        img_h = first_frame().pin_memory()
        img_d = img_h.to(dev)
        seed_ctr = [0]
        host_out = {}

        def to_host(name, t):
            # frames land in pinned host buffers (allocated once): the D2H is a real async DMA inside the timed region
            buf = host_out.get(name)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = host_out[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            return buf

        def step(img, fetch, use_shard=True):
            # every rank seeds its device generator identically: latents stay replicated over the rollout group
            seed_ctr[0] += 1
            noise = DeviceNoise(pipe, torch.Generator(device=dev).manual_seed(1234 + seed_ctr[0]))
            res = ro.generate(dict(img=img, prompt=list(ROLLOUT_PROMPTS), prompt_type="action"), noise=noise,
                              shard=sh if use_shard else None)
            if fetch:
                return to_host("img", res["pred_img"]), to_host("disp", res["pred_disparity"])
            return res["pred_img"], res["pred_disparity"]

        def resident():
            return img_d

        def from_host():
            return img_h.to(dev, non_blocking=True)

        h2d = nbytes(img_h)
    else:
        unit_h, dec_h = unit_inputs(dtype, pin=True)

        def to_dev(d, dec):
            e = dict(d)
            e["latents"] = d["latents"].to(dev, non_blocking=True)
            e["cond_tensors"] = [[c.to(dev, non_blocking=True) for c in st] for st in d["cond_tensors"]]
            e["block_noise"] = [c.to(dev, non_blocking=True) for c in d["block_noise"]]
            for k in ("enc", "pooled", "mask"):
                e[k] = d[k].to(dev, non_blocking=True)
            return e, [t.to(dev, non_blocking=True) for t in dec]

        def step(inp, fetch, use_shard=True):
            d, dec = inp
            s = sh if use_shard else None
            pipe.generate_one_unit(d["latents"], None, d["cond_tensors"], d["enc"], d["mask"], d["pooled"],
                                   STEPS_PER_STAGE, block_noise=d["block_noise"], shard=s)
            # the rollout decodes the first 16 (RGB) and next 16 (disparity) latent channels (pipeline.py:686-696)
            if s is not None:
                img, dsp = pipe.decode_latents_sharded([dec[0], dec[1]], s)
            else:
                img, dsp = pipe.decode_latent(dec[0]), pipe.decode_latent(dec[1])
            if fetch:
                return to_host("img", img), to_host("disp", dsp)
            return img, dsp

        res_in = to_dev(unit_h, dec_h)
        host_out = {}

        def to_host(name, t):
            buf = host_out.get(name)
            if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                buf = host_out[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            buf.copy_(t, non_blocking=True)
            return buf

        def resident():
            return res_in

        def from_host():
            return to_dev(unit_h, dec_h)

        h2d = sum(nbytes(unit_h[k]) for k in ("latents", "cond_tensors", "block_noise", "enc", "pooled", "mask")) + nbytes(dec_h)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    # ---- resident-input timing (value) -----------------------------------------------------------
    for _ in range(args.warmup):
        step(resident(), False)
    barrier()
    lib.dv_launch_count_reset()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(resident(), False)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = lib.dv_launch_count() // args.steps
    clk = clocks.stop()
    frames = frames_per_step(args.workload) * n_rollouts
    value = frames / (ms / 1e3)

    # ---- end-to-end through the public API with host buffers -------------------------------------
    e2e = None
    if not args.no_e2e:
        r = step(from_host(), True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            r = step(from_host(), True)
            torch.cuda.current_stream().synchronize()  # the frames are on the host before the next step
        e1.record()
        barrier()
        ms_e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        e2e = {"value": frames / (ms_e / 1e3), "unit": UNIT, "ms_per_step": ms_e, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": nbytes(r[0]) + nbytes(r[1]),
               "api": "B200Rollout.generate(batch_dict)" if rollout else "B200Pipeline.generate_one_unit + decode_latent"}

    # ---- N > 1: the same GPUs as N independent rollouts (replica throughput) -----------------------
    replicas = None
    if world > 1 and not args.no_replicas:
        step(resident(), False, use_shard=False)
        barrier()
        k = min(args.steps, 3)
        e0.record()
        for _ in range(k):
            step(resident(), False, use_shard=False)
        e1.record()
        barrier()
        ms_r = max_over_ranks(e0.elapsed_time(e1) / k)
        replicas = {"value": frames_per_step(args.workload) * world / (ms_r / 1e3), "unit": UNIT, "ms_per_step": ms_r,
                    "steps": k, "note": f"{world} independent single-GPU rollouts, no exchange (weak scaling)"}

    # ---- roofline: one extra profiled step, every launch bracketed by events -----------------------
    roofline = None
    if True:   # every rank runs the profiled step (a sharded rollout is collective); rank 0 reports its own launches
        lib.dv_profile_enable(1)
        lib.dv_profile_reset()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        step(resident(), False)
        p1.record()
        barrier()
        lib.dv_profile_enable(0)
        if rank == 0:
            path = args.profile_dump or os.path.join(tempfile.gettempdir(), f"dv_prof_{os.getpid()}.csv")
            _lib.check(lib.dv_profile_dump(path.encode()), "dv_profile_dump")
            peaks = {}
            try:
                peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            except Exception:
                pass
            roofline = roofline_from_profile(path, p0.elapsed_time(p1), peaks)
            roofline["whole_step"] = whole_step_summary(roofline, ms, gsz, peaks, rollout, args.layers)
            if not args.profile_dump:
                os.unlink(path)
        lib.dv_profile_reset()

    # ---- CPU baseline + same-box eager (rank 0, N = 1) -----------------------------------------------
    cpu, eager = None, None
    if rank == 0 and world == 1:
        if not args.no_same_box_eager:
            eager = same_box_eager(args)
        if not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            rs = ReferenceSampler(args.layers, "cpu", torch.float32)
            t = rs.step()
            sec = rs.scale(t, args.workload)
            cpu = {"value": frames_per_step(args.workload) / sec, "unit": UNIT, "cores": torch.get_num_threads(),
                   "kind": rs.kind, "sample": rs.describe(t)}

    if rank == 0:
        bad = [r_ for r_ in clk.get("reasons", []) if r_ in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong" if (world > 1 and n_rollouts == 1) else "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, parallelism_string(world, gsz)),
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "clock_rejected": bool(bad),
                "roofline": roofline, "cpu_baseline": cpu, "same_box_eager": eager, "replicas": replicas,
                "published_reference": {"value": 4.0, "unit": "frames/s", "hardware": "1x A800, full run.py pipeline",
                                        "source": "README.md:78", "comparable": False}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # libraries (NCCL banners, ...) may write to fd 1: keep the real stdout for the ONE JSON line
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
