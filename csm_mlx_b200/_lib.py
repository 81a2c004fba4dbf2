"""ctypes binding of libcsm_b200.so (the C ABI declared in include/csm_b200.h).

There is NO fallback: if the shared library is missing or the device is not sm_100 every product entry
point raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``csm_mlx_b200/csrc/build.sh``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

MAX_LAYERS = 16
PAGE = 16

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSMB_LIB_PATH") or os.path.join(_HERE, "libcsm_b200.so")  # override: A/B builds of the kernels


class CsmbError(RuntimeError):
    pass


class Sampler(C.Structure):
    _fields_ = [("temperature", C.c_float), ("top_k", C.c_int), ("top_p", C.c_float), ("min_p", C.c_float),
                ("min_keep", C.c_int), ("seed", C.c_uint64)]


class Llama(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("d_model", C.c_int), ("n_heads", C.c_int), ("n_kv_heads", C.c_int),
                ("head_dim", C.c_int), ("d_ff", C.c_int), ("eps", C.c_float),
                ("wqkv", C.c_void_p * MAX_LAYERS), ("wo", C.c_void_p * MAX_LAYERS),
                ("wgu", C.c_void_p * MAX_LAYERS), ("wdown", C.c_void_p * MAX_LAYERS),
                ("norm_in", C.c_void_p * MAX_LAYERS), ("norm_post", C.c_void_p * MAX_LAYERS),
                ("norm_final", C.c_void_p), ("rope", C.c_void_p)]


class Model(C.Structure):
    _fields_ = [("backbone", Llama), ("decoder", Llama), ("text_emb", C.c_void_p), ("audio_emb", C.c_void_p),
                ("projection", C.c_void_p), ("c0_head", C.c_void_p), ("audio_head_t", C.c_void_p),
                ("n_text_vocab", C.c_int), ("audio_vocab", C.c_int), ("n_codebooks", C.c_int),
                ("max_pos", C.c_int), ("weight_format", C.c_int), ("reserved", C.c_int)]


class Batch(C.Structure):
    _fields_ = [("batch", C.c_int), ("max_pages", C.c_int), ("kv_pool", C.c_void_p),
                ("kv_layer_stride", C.c_size_t), ("block_table", C.c_void_p), ("dec_kv_pool", C.c_void_p),
                ("dec_kv_layer_stride", C.c_size_t), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("flags", C.c_int)]


BATCH_ROW_INVARIANT = 1
WEIGHTS_BF16, WEIGHTS_E4M3 = 0, 1
ABI_VERSION = 3


class ChainOpts(C.Structure):
    """csmb_chain_opts: per-call switches of the fused batched chain (NULL = defaults)."""
    _fields_ = [("no_pdl", C.c_int), ("flags", C.c_int), ("smem_kb", C.c_int), ("proj_table", C.c_void_p)]


class FrameOpts(C.Structure):
    """csmb_frame_opts: per-call switches of the batch-1 frame kernel (NULL = defaults)."""
    _fields_ = [("ctas", C.c_int), ("flags", C.c_int), ("prefetch_stages", C.c_int), ("prefetch_interval", C.c_int),
                ("prof", C.c_void_p)]


class Tc3(C.Structure):
    """csmb_tc3: one tensor-core strided-row GEMM of the codec (see include/csm_b200.h)."""
    _fields_ = [("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("a_rows", C.c_longlong), ("lda", C.c_int), ("rpb", C.c_int),
                ("C", C.c_int), ("taps", C.c_int), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_rows", C.c_int),
                ("ldw", C.c_int), ("y32", C.c_void_p), ("y_batch", C.c_longlong), ("ldy", C.c_int), ("y_hi", C.c_void_p),
                ("y_lo", C.c_void_p), ("p_batch", C.c_longlong), ("ldp", C.c_int), ("plane_act", C.c_int),
                ("bias", C.c_void_p), ("scale", C.c_void_p), ("residual", C.c_void_p), ("r_batch", C.c_longlong),
                ("ldr", C.c_int), ("B", C.c_int), ("T", C.c_int), ("N", C.c_int), ("act_out", C.c_int),
                ("err_flag", C.c_void_p)]


_lib: Optional[C.CDLL] = None

_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_SIGS = {
    "csmb_abi_version": (C.c_int, []),
    "csmb_strerror": (C.c_char_p, [_I]),
    "csmb_last_cuda_error": (C.c_char_p, []),
    "csmb_check_device": (C.c_int, [_I]),
    "csmb_debug_launch_count": (C.c_ulonglong, []),
    "csmb_embed_sum": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "csmb_embed_audio": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_rmsnorm": (C.c_int, [_P, _I, _P, _P, _I, _I, _I, C.c_float, _I, _P]),
    "csmb_linear": (C.c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_e4m3_blob_bytes": (C.c_size_t, [_I, _I]),
    "csmb_linear_e4m3": (C.c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_linear_tc_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "csmb_linear_tc": (C.c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _I, _P, C.c_size_t, _I, _P]),
    "csmb_swiglu": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "csmb_rope_kv_append": (C.c_int, [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P]),
    "csmb_attention": (C.c_int, [_P, _I, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "csmb_sample": (C.c_int, [_P, _I, _P, _I, _I, _I, C.POINTER(Sampler), C.c_uint64, _P, C.c_uint32, _I, _P]),
    "csmb_lm_workspace_bytes": (C.c_size_t, [C.POINTER(Model), _I]),
    "csmb_backbone_forward": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), _P, _P, _P, _P, _I, _P, _I, _P, _P,
                                        _I, _P]),
    "csmb_depth_decode": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), _P, _P, C.POINTER(Sampler), C.c_uint64,
                                    _P, _P, _P, _I, _I, _I, _P]),
    "csmb_decode_frame": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), _P, _P, _P, C.POINTER(Sampler),
                                    C.c_uint64, _I, _P]),
    "csmb_decode_frame_fast_workspace_bytes": (C.c_size_t, [C.POINTER(Model), _I]),
    "csmb_decode_frame_fast_supported": (C.c_int, [C.POINTER(Model), C.POINTER(Sampler)]),
    "csmb_decode_frame_fast": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), _P, _P, _P, C.POINTER(Sampler),
                                         C.c_uint64, _P, C.c_size_t, _I, _P]),
    "csmb_decode_frame_fast_admit": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), _P, _P, _P, C.POINTER(Sampler),
                                               C.c_uint64, _P, _P, C.POINTER(ChainOpts), _P, C.c_size_t, _I, _P]),
    "csmb_prefill_fast_workspace_bytes": (C.c_size_t, [C.POINTER(Model), _I]),
    "csmb_prefill_fast": (C.c_int, [C.POINTER(Model), C.POINTER(Batch), _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, C.c_size_t,
                                    _I, _P]),
    "csmb_proj_table_bytes": (C.c_size_t, [C.POINTER(Model)]),
    "csmb_proj_table_workspace_bytes": (C.c_size_t, [C.POINTER(Model)]),
    "csmb_build_proj_table": (C.c_int, [C.POINTER(Model), _P, _P, C.c_size_t, _I, _P]),
    "csmb_frame_workspace_bytes": (C.c_size_t, [C.POINTER(Model), _I]),
    "csmb_frame_b1": (C.c_int, [C.POINTER(Model), _P, C.c_size_t, _P, _P, _P, _P, C.POINTER(Sampler), C.c_uint64,
                                C.POINTER(FrameOpts), _P, C.c_size_t, _P, _I, _P]),
    "csmb_frame_b1_slot": (C.c_int, [C.POINTER(Model), _P, C.c_size_t, _P, _P, _P, _P, C.POINTER(Sampler), C.c_uint64, _I,
                                     C.POINTER(FrameOpts), _P, C.c_size_t, _P, _I, _P]),
    "csmb_frame_b1_depth": (C.c_int, [C.POINTER(Model), _P, _P, _P, C.POINTER(Sampler), C.c_uint64, C.POINTER(FrameOpts), _P,
                                      C.c_size_t, _P, _I, _P]),
    "csmb_gemm_f32": (C.c_int, [_P, _LL, _I, _P, _P, _LL, _I, _P, _P, _P, _LL, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_gemm_tc3": (C.c_int, [C.POINTER(Tc3), _I, _P]),
    "csmb_split_planes": (C.c_int, [_P, _LL, _I, _P, _P, _LL, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_layernorm_planes": (C.c_int, [_P, _LL, _P, _P, _P, _P, _I, _I, _I, C.c_float, _I, _P]),
    "csmb_mimi_attention_planes": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "csmb_conv_in_planes": (C.c_int, [_P, _LL, _P, _P, _P, _P, _P, _LL, _I, _I, _I, _I, _I, _P]),
    "csmb_rvq_argmin_update_planes": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_layernorm": (C.c_int, [_P, _LL, _P, _P, _P, _I, _I, _I, C.c_float, _I, _P]),
    "csmb_mimi_attention": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_rvq_gather": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_upsample_dw": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "csmb_rvq_argmin_update": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_copy_rows": (C.c_int, [_P, _LL, _I, _P, _LL, _I, _I, _I, _I, _I, _I, _P]),
    "csmb_shift_rows": (C.c_int, [_P, _LL, _I, _I, _I, _I, _I, _P]),
    "csmb_add_int": (C.c_int, [_P, _I, _I, _P]),
}


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_SIGS)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CsmbError(f"{LIB_PATH} not found: the CUDA library is not built and there is no fallback path")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.csmb_abi_version() != ABI_VERSION:
            raise CsmbError(f"{LIB_PATH} has ABI version {handle.csmb_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        l = lib()
        msg = l.csmb_strerror(status).decode()
        if status == -2:
            msg += ": " + l.csmb_last_cuda_error().decode()
        raise CsmbError(f"libcsm_b200: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise CsmbError("libcsm_b200 takes device tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise CsmbError("non-contiguous tensor passed to libcsm_b200")
    return t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_device(device: torch.device) -> int:
    """Returns the CUDA ordinal; raises if there is no sm_100 GPU (no CPU path exists)."""
    if device.type != "cuda":
        raise CsmbError("csm_mlx_b200 runs on an sm_100a (B200) GPU only; there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(lib().csmb_check_device(idx))
    return idx
