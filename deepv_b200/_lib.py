"""ctypes binding of libdeepv_b200.so (the C ABI declared in include/deepv_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdeepv_b200.so"

DV_DTYPE_F32 = 0
DV_DTYPE_BF16 = 1

_lib = None


class DeepVError(RuntimeError):
    pass


class MMDiTConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "num_layers", "num_heads", "head_dim", "in_channels", "patch_size", "joint_dim",
        "pooled_dim", "pos_embed_max", "pos_base_size", "patch_k_pad")]


_PP = C.POINTER(C.c_void_p)


class MMDiTWeights(C.Structure):
    _fields_ = [
        ("w_qkv_x", _PP), ("b_qkv_x", _PP), ("w_qkv_c", _PP), ("b_qkv_c", _PP),
        ("qk_norm_x", _PP), ("qk_norm_c", _PP),
        ("w_out_x", _PP), ("b_out_x", _PP), ("w_out_c", _PP), ("b_out_c", _PP),
        ("w_ff1_x", _PP), ("b_ff1_x", _PP), ("w_ff2_x", _PP), ("b_ff2_x", _PP),
        ("w_ff1_c", _PP), ("b_ff1_c", _PP), ("w_ff2_c", _PP), ("b_ff2_c", _PP),
        ("w_mod", C.c_void_p), ("b_mod", C.c_void_p), ("mod_rows", C.c_int),
        ("w_t1", C.c_void_p), ("b_t1", C.c_void_p), ("w_t2", C.c_void_p), ("b_t2", C.c_void_p),
        ("w_p1", C.c_void_p), ("b_p1", C.c_void_p), ("w_p2", C.c_void_p), ("b_p2", C.c_void_p),
        ("w_ctx", C.c_void_p), ("b_ctx", C.c_void_p),
        ("w_patch", C.c_void_p), ("b_patch", C.c_void_p),
        ("w_patch_hist", C.c_void_p), ("b_patch_hist", C.c_void_p),
        ("w_proj_out", C.c_void_p), ("b_proj_out", C.c_void_p),
    ]


class VAEConfig(C.Structure):
    _fields_ = [
        ("latent_channels", C.c_int), ("out_channels", C.c_int),
        ("block_channels", C.c_int * 4), ("layers_per_block", C.c_int * 4),
        ("spatial_up", C.c_int * 4), ("temporal_up", C.c_int * 4), ("norm_groups", C.c_int),
        ("enc_in_channels", C.c_int), ("enc_block_channels", C.c_int * 4),
        ("enc_layers_per_block", C.c_int * 4), ("enc_spatial_down", C.c_int * 4),
        ("enc_temporal_down", C.c_int * 4),
    ]


class TensorRef(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("numel", C.c_longlong)]


# every symbol include/deepv_b200.h declares: name -> (restype, argtypes)
_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
SIGNATURES = {
    "dv_last_error": (C.c_char_p, []),
    "dv_version": (_i, []),
    "dv_launch_count": (_ll, []),
    "dv_launch_count_reset": (None, []),
    "dv_profile_enable": (None, [_i]),
    "dv_profile_reset": (None, []),
    "dv_profile_dump": (_i, [C.c_char_p]),
    "dv_profile_summary": (_i, [_i, C.POINTER(_ll), C.POINTER(_d), C.POINTER(_d), C.POINTER(_d)]),
    "dv_cfg_euler_step": (_i, [_vp, _i, _vp, _vp, _ll, _f, _f, _d, _d, _i, _vp]),
    "dv_stage_renoise": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _d, _i, _vp]),
    "dv_block_noise": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp]),
    "dv_resize_half": (_i, [_vp, _vp, _ll, _i, _i, _f, _i, _vp]),
    "dv_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dv_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dv_conv3d_cl": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dv_conv3d_strided_cl": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dv_mmdit_create": (_i, [C.POINTER(MMDiTConfig), C.POINTER(MMDiTWeights), C.POINTER(_vp)]),
    "dv_mmdit_destroy": (None, [_vp]),
    "dv_mmdit_debug_buffer": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_ll)]),
    "dv_mmdit_plan_create": (_i, [_vp, _i, _i, C.POINTER(_i), _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "dv_mmdit_plan_destroy": (None, [_vp]),
    "dv_mmdit_plan_workspace_bytes": (_ll, [_vp]),
    "dv_mmdit_plan_flops": (_d, [_vp]),
    "dv_mmdit_forward": (_i, [_vp, _PP, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "dv_mmdit_plan_set_sp": (_i, [_vp, _i, _i, _vp, _vp]),
    "dv_mmdit_plan_buffers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "dv_mmdit_plan_set_sp_peers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "dv_ipc_get_handle": (_i, [_vp, _vp]),
    "dv_ipc_open_handle": (_i, [_vp, C.POINTER(_vp)]),
    "dv_ipc_close_handle": (_i, [_vp]),
    "dv_comm_unique_id": (_i, [C.c_char_p, _vp]),
    "dv_comm_create": (_i, [C.c_char_p, _vp, _i, _i, C.POINTER(_vp)]),
    "dv_comm_destroy": (None, [_vp]),
    "dv_comm_exchange": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "dv_vae_create": (_i, [C.POINTER(VAEConfig), C.POINTER(TensorRef), _i, C.POINTER(_vp)]),
    "dv_vae_destroy": (None, [_vp]),
    "dv_vae_plan_create": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "dv_vae_plan_destroy": (None, [_vp]),
    "dv_vae_plan_flops": (_d, [_vp]),
    "dv_vae_plan_set_first_frame": (_i, [_vp, _i]),
    "dv_vae_decode": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "dv_vae_enc_plan_create": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "dv_vae_enc_plan_destroy": (None, [_vp]),
    "dv_vae_enc_plan_flops": (_d, [_vp]),
    "dv_vae_enc_plan_latent_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dv_vae_encode": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    "dv_gaussian_sample": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "dv_vae_plan_geometry": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "dv_vae_plan_tile_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "dv_vae_plan_bind_tile": (_i, [_vp, _i, _vp]),
    "dv_vae_decode_tiles": (_i, [_vp, _vp, _i, C.c_ulonglong, _vp]),
    "dv_vae_blend": (_i, [_vp, _vp, _i, _vp]),
    "dv_frames_requantise": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "dv_disparity_post": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dv_disparity_renorm": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp]),
    "dv_raymap_to_pose": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dv_camera_raymap": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
}


def load() -> C.CDLL:
    """Load the shared library (building it is `python -m deepv_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise DeepVError(
            f"{LIB_PATH} is missing: run `python -m deepv_b200.build` (nvcc, sm_100a). "
            "deepv_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dv_last_error().decode(errors="replace")
        raise DeepVError(f"{what or 'deepv_b200 call'} failed ({rc}): {msg}")


def dtype_code(t) -> int:
    import torch
    if t == torch.float32:
        return DV_DTYPE_F32
    if t == torch.bfloat16:
        return DV_DTYPE_BF16
    raise DeepVError(f"unsupported dtype {t}; the sm_100a path takes float32 or bfloat16")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DeepVError(
                "deepv_b200 kernels need CUDA tensors on a B200 (sm_100a); got a CPU tensor. "
                "There is no CPU fallback — use the oracle/ only as a test checker.")
