"""Host-side mirror of the reference denoiser (`model/mmdit.py:MMDiT`) on the sm_100a kernels.

`B200MMDiT` is what `InferencePipeline._create_models` (pipeline.py:203-223) returns in place
of the reference `MMDiT`: same attributes the pipeline touches (`in_channels`, `.eval()`,
`.to()`), same keyword-only call (`pipeline.py:487-497`), same return value (a list whose [0]
is `[B, C, 1, h, w]`).  All arithmetic runs inside libdeepv_b200.so (dv_mmdit_forward); torch
only owns device memory and the stream.  No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import MMDiTConfig, MMDiTWeights, check


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


class B200MMDiT:
    """Drop-in for reference `MMDiT` (inference only)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], config: dict, device="cuda",
                 out_dtype: Optional[torch.dtype] = None):
        self.lib = _lib.load()
        self.config = dict(config)
        self.in_channels = config["in_channels"]           # pipeline.py:551
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DeepVError("B200MMDiT needs a CUDA device (sm_100a); no CPU fallback")
        self.out_dtype = out_dtype
        self._keep: List[torch.Tensor] = []
        self._plans: Dict[tuple, int] = {}
        self._handle = C.c_void_p()
        self._pack(state_dict)

    # -- reference-compatible no-ops -------------------------------------------------------
    def eval(self):
        return self

    def to(self, *args, **kwargs):
        return self

    # -- weight packing ------------------------------------------------------------------------
    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(t)
        return t

    def _pack(self, sd: Dict[str, torch.Tensor]) -> None:
        cfg = self.config
        NL = cfg["num_layers"]
        H, hd = cfg["num_attention_heads"], cfg["attention_head_dim"]
        D = H * hd
        Cc, P = cfg["in_channels"], cfg["patch_size"]
        kpad = (Cc * P * P + 63) // 64 * 64
        bf, f32 = torch.bfloat16, torch.float32
        per: Dict[str, list] = {k: [] for k, _ in MMDiTWeights._fields_ if k.startswith(("w_", "b_", "qk_"))
                                and k.split("_")[-1] in ("x", "c")}

        def cat(names):
            return torch.cat([sd[n] for n in names], dim=0)

        mod_w, mod_b = [], []
        for i in range(NL):
            b = f"transformer_blocks.{i}."
            a = b + "attn."
            last = i == NL - 1
            per["w_qkv_x"].append(self._dev(cat([a + "to_q.weight", a + "to_k.weight", a + "to_v.weight"]), bf))
            per["b_qkv_x"].append(self._dev(cat([a + "to_q.bias", a + "to_k.bias", a + "to_v.bias"]), f32))
            per["w_qkv_c"].append(self._dev(cat([a + "add_q_proj.weight", a + "add_k_proj.weight", a + "add_v_proj.weight"]), bf))
            per["b_qkv_c"].append(self._dev(cat([a + "add_q_proj.bias", a + "add_k_proj.bias", a + "add_v_proj.bias"]), f32))
            per["qk_norm_x"].append(self._dev(cat([a + "norm_q.weight", a + "norm_k.weight"]), f32))
            per["qk_norm_c"].append(self._dev(cat([a + "norm_add_q.weight", a + "norm_add_k.weight"]), f32))
            per["w_out_x"].append(self._dev(sd[a + "to_out.0.weight"], bf))
            per["b_out_x"].append(self._dev(sd[a + "to_out.0.bias"], f32))
            per["w_ff1_x"].append(self._dev(sd[b + "ff.net.0.proj.weight"], bf))
            per["b_ff1_x"].append(self._dev(sd[b + "ff.net.0.proj.bias"], f32))
            per["w_ff2_x"].append(self._dev(sd[b + "ff.net.2.weight"], bf))
            per["b_ff2_x"].append(self._dev(sd[b + "ff.net.2.bias"], f32))
            if last:
                for k in ("w_out_c", "b_out_c", "w_ff1_c", "b_ff1_c", "w_ff2_c", "b_ff2_c"):
                    per[k].append(None)
            else:
                per["w_out_c"].append(self._dev(sd[a + "to_add_out.weight"], bf))
                per["b_out_c"].append(self._dev(sd[a + "to_add_out.bias"], f32))
                per["w_ff1_c"].append(self._dev(sd[b + "ff_context.net.0.proj.weight"], bf))
                per["b_ff1_c"].append(self._dev(sd[b + "ff_context.net.0.proj.bias"], f32))
                per["w_ff2_c"].append(self._dev(sd[b + "ff_context.net.2.weight"], bf))
                per["b_ff2_c"].append(self._dev(sd[b + "ff_context.net.2.bias"], f32))
            mod_w += [sd[b + "norm1.linear.weight"], sd[b + "norm1_context.linear.weight"]]
            mod_b += [sd[b + "norm1.linear.bias"], sd[b + "norm1_context.linear.bias"]]
        mod_w.append(sd["norm_out.linear.weight"])
        mod_b.append(sd["norm_out.linear.bias"])

        w = MMDiTWeights()
        self._arrays = {}
        for k, lst in per.items():
            arr = _ptr_array(lst)
            self._arrays[k] = arr
            setattr(w, k, C.cast(arr, C.POINTER(C.c_void_p)))
        wm = self._dev(torch.cat(mod_w, dim=0), bf)
        w.w_mod, w.b_mod, w.mod_rows = wm.data_ptr(), self._dev(torch.cat(mod_b), f32).data_ptr(), wm.shape[0]

        def put(wname, bname, wt, bt):
            setattr(w, wname, self._dev(wt, bf).data_ptr())
            setattr(w, bname, self._dev(bt, f32).data_ptr())

        te = "time_text_embed."
        put("w_t1", "b_t1", sd[te + "timestep_embedder.linear_1.weight"], sd[te + "timestep_embedder.linear_1.bias"])
        put("w_t2", "b_t2", sd[te + "timestep_embedder.linear_2.weight"], sd[te + "timestep_embedder.linear_2.bias"])
        put("w_p1", "b_p1", sd[te + "text_embedder.linear_1.weight"], sd[te + "text_embedder.linear_1.bias"])
        put("w_p2", "b_p2", sd[te + "text_embedder.linear_2.weight"], sd[te + "text_embedder.linear_2.bias"])
        put("w_ctx", "b_ctx", sd["context_embedder.weight"], sd["context_embedder.bias"])

        def patch_w(name):  # Conv2d [D, C, p, p] -> [D, kpad], k = c*p*p + p1*p + p2
            m = sd[name].reshape(D, Cc * P * P)
            return torch.nn.functional.pad(m, (0, kpad - m.shape[1]))

        put("w_patch", "b_patch", patch_w("pos_embed.proj.weight"), sd["pos_embed.proj.bias"])
        put("w_patch_hist", "b_patch_hist", patch_w("pos_embed.proj_history.weight"), sd["pos_embed.proj_history.bias"])
        nout = sd["proj_out.bias"].shape[0]
        bpad = torch.nn.functional.pad(sd["proj_out.bias"], (0, (nout + 31) // 32 * 32 - nout))
        put("w_proj_out", "b_proj_out", sd["proj_out.weight"], bpad)

        c = MMDiTConfig(NL, H, hd, Cc, P, cfg["joint_attention_dim"], cfg["pooled_projection_dim"],
                        cfg["pos_embed_max_size"], cfg["sample_size"] // P, kpad)
        self._weights = w
        check(self.lib.dv_mmdit_create(C.byref(c), C.byref(w), C.byref(self._handle)), "dv_mmdit_create")

    # -- Ulysses sequence parallelism (no reference counterpart; BASELINE.json north_star) --------
    def set_sequence_parallel(self, sp_rank: int, sp_world: int, exchange=None, user=None):
        """Shard the video tokens of every forward over `sp_world` ranks (this one is `sp_rank`).
        `exchange` is the equal-block all-to-all the forward calls around every attention: either
        a `parallel.NcclExchange` (NCCL inside the library, no Python in the loop) or a Python
        callable `(send_ptr, recv_ptr, bytes_per_peer, stream_ptr) -> int` (tests)."""
        if sp_world > 1 and exchange is None:
            raise _lib.DeepVError("set_sequence_parallel: an exchange is required for sp_world > 1")
        self._sp = (int(sp_rank), int(sp_world))
        self._sp_keep = exchange
        if exchange is None:
            self._sp_fn, self._sp_user = None, None
        elif hasattr(exchange, "fn_ptr"):
            self._sp_fn, self._sp_user = exchange.fn_ptr, exchange.user_ptr
        else:
            proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p)

            def tramp(_user, send, recv, nbytes, stream):
                try:
                    return int(exchange(send, recv, nbytes, stream) or 0)
                except Exception as e:  # noqa: BLE001 - must not unwind through C
                    import traceback
                    traceback.print_exc()
                    return -1
            self._sp_cb = proto(tramp)
            self._sp_fn, self._sp_user = C.cast(self._sp_cb, C.c_void_p), None
        # applied lazily per plan (a setting may not divide the token count of every cached layout)

    # -- plans -------------------------------------------------------------------------------------
    def _plan(self, B: int, clip_dims: tuple, text_len: int, hist: Optional[tuple], hist_ds: int):
        key = (B, clip_dims, text_len, hist, hist_ds)
        p = self._plans.get(key)
        if p is None:
            flat = [v for d in clip_dims for v in d]
            arr = (C.c_int * len(flat))(*flat)
            h = C.c_void_p()
            check(self.lib.dv_mmdit_plan_create(self._handle, B, len(clip_dims), arr, text_len,
                                                1 if hist else 0, hist[0] if hist else 0,
                                                hist[1] if hist else 0, hist_ds or 0, C.byref(h)),
                  "dv_mmdit_plan_create")
            p = h.value
            self._plans[key] = p
        sp = getattr(self, "_sp", (0, 1))
        applied = getattr(self, "_plan_sp", None)
        if applied is None:
            applied = self._plan_sp = {}
        want = (sp[0], sp[1], id(getattr(self, "_sp_keep", None)))
        if applied.get(p, (0, 1, id(None))) != want:
            fn = getattr(self, "_sp_fn", None) if sp[1] > 1 else None
            user = getattr(self, "_sp_user", None) if sp[1] > 1 else None
            check(self.lib.dv_mmdit_plan_set_sp(p, sp[0], sp[1], fn, user), "dv_mmdit_plan_set_sp")
            keep = getattr(self, "_sp_keep", None)
            if sp[1] > 1 and callable(getattr(keep, "peer_pointers", None)):
                # peer-memory exchange: every rank of the group maps the others' qkv / attention buffers, fp32 stream
                # and barrier flag words (collective: all ranks of the SP group create their plans in the same order)
                mine = [C.c_void_p() for _ in range(4)]
                check(self.lib.dv_mmdit_plan_buffers(p, *[C.byref(m) for m in mine]), "dv_mmdit_plan_buffers")
                lists = keep.peer_pointers(key, [m.value for m in mine])
                arrs = [(C.c_void_p * sp[1])(*lst) for lst in lists]
                if getattr(keep, "device_barrier", True) and len(arrs) == 4:
                    check(self.lib.dv_mmdit_plan_set_sp_peers(p, *arrs), "dv_mmdit_plan_set_sp_peers")
                else:   # NCCL (or the host callback) stays the barrier and carries the final all-gather
                    check(self.lib.dv_mmdit_plan_set_sp_peers(p, arrs[0], arrs[1], None, None), "dv_mmdit_plan_set_sp_peers")
            else:
                check(self.lib.dv_mmdit_plan_set_sp_peers(p, None, None, None, None), "dv_mmdit_plan_set_sp_peers")
            applied[p] = want
        return p

    def plan_flops(self, B, clip_dims, text_len=77, hist=None, hist_ds=2) -> float:
        return self.lib.dv_mmdit_plan_flops(self._plan(B, tuple(clip_dims), text_len, hist, hist_ds))

    # -- forward (reference signature: mmdit.py:1467-1478) ----------------------------------------
    def __call__(self, sample=None, encoder_hidden_states=None, encoder_attention_mask=None,
                 pooled_projections=None, timestep_ratio=None, history=None, history_mask=None,
                 history_downsample_ratio=None):
        if len(sample) != 1:
            raise _lib.DeepVError("B200MMDiT supports one stage list per call (pipeline.py:488)")
        clips = sample[0] if isinstance(sample[0], (list, tuple)) else [sample[0]]
        _lib.require_cuda(*clips, encoder_hidden_states, pooled_projections, timestep_ratio, history)
        io_dtype = clips[-1].dtype
        clips = [c.to(io_dtype).contiguous() for c in clips]
        B = clips[-1].shape[0]
        dims = tuple((c.shape[2], c.shape[3], c.shape[4]) for c in clips)
        enc = encoder_hidden_states
        if enc.dtype not in (torch.float32, torch.bfloat16):
            enc = enc.float()
        enc = enc.contiguous()
        text_len = enc.shape[1]
        masks = [encoder_attention_mask.to(device=self.device, dtype=torch.float32)]
        hist_hw = None
        if history is not None:
            history = history.to(io_dtype).contiguous()
            hist_hw = (history.shape[-2], history.shape[-1])
            masks.insert(0, history_mask.to(device=self.device, dtype=torch.float32))
        ctx_mask = torch.cat(masks, dim=1).contiguous()
        pooled = pooled_projections.to(torch.float32).contiguous()
        # the timestep must stay fp32 when judged against the fp32 oracle;
        # whatever dtype the caller rounded it to is what gets embedded, like the reference
        t = timestep_ratio.to(torch.float32).contiguous()
        plan = self._plan(B, dims, text_len, hist_hw, history_downsample_ratio or 0)
        out_dtype = self.out_dtype or io_dtype
        _, Cc, _, h, w = clips[-1].shape
        out = torch.empty((B, Cc, 1, h, w), device=self.device, dtype=out_dtype)
        ptrs = _ptr_array(clips)
        check(self.lib.dv_mmdit_forward(plan, C.cast(ptrs, C.POINTER(C.c_void_p)),
                                        _lib.dtype_code(io_dtype), enc.data_ptr(),
                                        _lib.dtype_code(enc.dtype), ctx_mask.data_ptr(),
                                        pooled.data_ptr(), t.data_ptr(),
                                        history.data_ptr() if history is not None else None,
                                        out.data_ptr(), _lib.dtype_code(out_dtype), _lib.stream_ptr()),
              "dv_mmdit_forward")
        # keep the staging tensors alive until the stream has consumed them
        self._last_inputs = (clips, enc, ctx_mask, pooled, t, history)
        return [out]

    forward = __call__

    def close(self):
        for p in self._plans.values():
            self.lib.dv_mmdit_plan_destroy(p)
        self._plans.clear()
        if self._handle:
            self.lib.dv_mmdit_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
