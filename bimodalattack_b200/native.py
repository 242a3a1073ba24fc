"""ctypes binding of ``libbimodal_b200.so`` (the C ABI declared in ``include/bimodal_b200.h``).

PyTorch is used only for device memory and streams: every call below passes raw device pointers
(``tensor.data_ptr()``) and the current CUDA stream handle. There is no CPU or eager fallback: if
the shared library is missing, or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_LIB_NAME = "libbimodal_b200.so"
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)

DTYPE_CODES = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


class ModelDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int), ("d_model", C.c_int), ("n_heads", C.c_int),
        ("n_kv_heads", C.c_int), ("head_dim", C.c_int), ("d_ff", C.c_int), ("vocab", C.c_int),
        ("rms_eps", C.c_float), ("norm_mode", C.c_int), ("act", C.c_int),
        ("embed_scale", C.c_float), ("attn_scale", C.c_float),
        ("max_rows", C.c_int), ("max_prefix", C.c_int),
        ("qk_norm", C.c_int), ("sandwich_norm", C.c_int),
    ]


class LayerWeights(C.Structure):
    _fields_ = [
        ("w_qkv", C.c_void_p), ("w_o", C.c_void_p), ("w_gate_up", C.c_void_p),
        ("w_down", C.c_void_p), ("norm_attn", C.c_void_p), ("norm_mlp", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("norm_post_attn", C.c_void_p), ("norm_post_mlp", C.c_void_p),
        ("q_norm", C.c_void_p), ("k_norm", C.c_void_p),
    ]


class GemmDesc(C.Structure):
    """``bma_gemm_desc`` (include/bimodal_b200.h)."""
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong), ("B", C.c_void_p), ("ldb", C.c_longlong),
        ("C", C.c_void_p), ("C32", C.c_void_p), ("ldc", C.c_longlong),
        ("R", C.c_void_p), ("ldr", C.c_longlong), ("bias", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("block_n", C.c_int), ("ctas", C.c_int),
        ("batch", C.c_int), ("a_row", C.c_int), ("a_col", C.c_int), ("b_row", C.c_int), ("b_col", C.c_int),
        ("c_batch", C.c_longlong), ("transpose_out", C.c_int), ("causal", C.c_int),
    ]


# name -> (restype, argtypes); mirrors include/bimodal_b200.h one to one
_P, _LL, _I, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_float
_ULL, _U32 = C.c_ulonglong, C.c_uint
SIGNATURES = {
    "bma_version": (_I, []),
    "bma_last_error": (C.c_char_p, []),
    "bma_watchdog_code": (_I, [C.POINTER(_U32)]),
    "bma_launch_count": (_ULL, []),
    "bma_pgd_step": (_I, [_P, _P, _P, _LL, _F, _F, _P, C.POINTER(_F), C.POINTER(_F), _I, _LL, _P]),
    "bma_token_grad": (_I, [_P, _LL, _P, _LL, _P, _LL, _I, _I, _I, _P]),
    "bma_sampler_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "bma_sample_ids": (_I, [_P, _P, _I, _LL, _I, _I, _P, _I, _I, _I, _ULL, _ULL, _P, _P, _P,
                            C.POINTER(_ULL), _P]),
    "bma_scorer_create": (_P, [C.POINTER(ModelDesc)]),
    "bma_scorer_destroy": (None, [_P]),
    "bma_scorer_set_layer": (_I, [_P, _I, C.POINTER(LayerWeights)]),
    "bma_scorer_set_head": (_I, [_P, _P, _P, _P, _LL]),
    "bma_scorer_workspace_bytes": (C.c_size_t, [_P]),
    "bma_scorer_profile": (_I, [_P, _I, _I]),
    "bma_scorer_profile_read": (_I, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_LL), _I]),
    "bma_scorer_prefill": (_I, [_P, _P, _I, _P]),
    "bma_scorer_prefill_deferred": (_I, [_P, _P, _I]),
    "bma_scorer_reset_key": (_I, [_P, _P, _P, _P]),
    "bma_scorer_score": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _I, _I, _P, _P, _P, _U32, _P, _I, _P]),
    "bma_scorer_group_score": (_I, [C.POINTER(_P), _I, _P, _I, _I, _P, _P, _I, _P, _I, _I, _P, _P, _P, _P, _I, _I, _P]),
    "bma_scorer_export_ipc": (_I, [_P, _P]),
    "bma_scorer_import_ipc": (_I, [_P, _P, _I, _I]),
    "bma_scorer_clear_peers": (_I, [_P]),
    "bma_scorer_exchange_wait_ns": (_I, [_P, C.POINTER(_ULL), _I]),
    "bma_scorer_peers_barrier": (_I, [_P, _I, _P]),
    "bma_scorer_set_early_signal": (_I, [_P, _I]),
    "bma_scorer_set_comm": (_I, [_P, _P]),
    "bma_comm_unique_id": (_I, [_P]),
    "bma_comm_create": (_P, [_P, _I, _I]),
    "bma_comm_destroy": (None, [_P]),
    "bma_comm_rank": (_I, [_P]),
    "bma_comm_size": (_I, [_P]),
    "bma_comm_last_error": (C.c_char_p, []),
    "bma_comm_allreduce_min_packed": (_I, [_P, _P, _I, _P]),
    "bma_comm_reduce_winner": (_I, [_P, _P, _P, _P]),
    "bma_scorer_last_rows": (_LL, [_P]),
    "bma_scorer_set_dedupe": (_I, [_P, _I]),
    "bma_scorer_set_fused_publish": (_I, [_P, _I]),
    "bma_scorer_score_host": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _I, _I, _P, _P, _U32, _P, _I, _P]),
    "bma_gemm_ex": (_I, [C.POINTER(GemmDesc), _P]),
    "bma_qkv_rope": (_I, [_P, _LL, _P, _P, _LL, _P, _LL, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "bma_softmax_rows": (_I, [_P, _LL, _P, _LL, _I, _I, _F, _I, _P]),
    "bma_softmax_bwd_rows": (_I, [_P, _LL, _P, _LL, _P, _LL, _I, _I, _F, _I, _P]),
    "bma_transpose_bf16": (_I, [_P, _LL, _LL, _P, _LL, _LL, _I, _I, _I, _P]),
    "bma_rmsnorm_bwd": (_I, [_P, _LL, _P, _LL, _P, _P, _LL, _P, _LL, _I, _I, _F, _I, _P]),
    "bma_layernorm": (_I, [_P, _LL, _P, _P, _P, _LL, _I, _I, _F, _P]),
    "bma_layernorm_bwd": (_I, [_P, _LL, _P, _LL, _P, _P, _LL, _P, _LL, _I, _I, _F, _P]),
    "bma_rope_bwd": (_I, [_P, _LL, _P, _P, _I, _I, _I, _I, _P]),
    "bma_glu": (_I, [_P, _LL, _P, _LL, _I, _I, _I, _P]),
    "bma_glu_bwd": (_I, [_P, _LL, _P, _LL, _P, _LL, _I, _I, _I, _P]),
    "bma_act": (_I, [_P, _P, _P, _LL, _I, _P]),
    "bma_ce_grad": (_I, [_P, _LL, _P, _P, _LL, _P, _I, _I, _P]),
    "bma_patchify": (_I, [_P, _P, _LL, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _P]),
    "bma_unpatchify_grad": (_I, [_P, _LL, _P, _I, _I, _I, _I, C.POINTER(_F), _P]),
    "bma_add_rows": (_I, [_P, _LL, _P, _LL, _I, _P, _LL, _I, _I, _P]),
    "bma_gemm_bf16": (_I, [_P, _LL, _P, _LL, _P, _LL, _P, _LL, _I, _I, _I, _I, _P]),
    "bma_rmsnorm": (_I, [_P, _LL, _P, _P, _LL, _I, _I, _F, _I, _P]),
    "bma_add_norm": (_I, [_P, _LL, _P, _P, _LL, _P, _P, _LL, _I, _I, _F, _I, _P]),
    "bma_attention": (_I, [_P, _LL, _P, _LL, _P, _LL, _P, _LL, _P, _LL, _P, _LL, _I, _I, _I, _I,
                           _I, _I, _F, _P]),
}

_lib: Optional[C.CDLL] = None


def library_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_NAME} not found at {_LIB_PATH}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or bimodalattack_b200/csrc/build.sh). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().bma_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"bimodal_b200 {what} failed (cudaError {rc}): {msg}")


def stream_ptr(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"bimodal_b200: `{name}` must be a CUDA tensor (got {t.device}); this package has no CPU path"
        )


# ------------------------------------------------------------------------------------------------
# thin functional wrappers (tensors in, tensors out)

def pgd_step_(image: torch.Tensor, grad: torch.Tensor, orig: torch.Tensor, step: float, eps: float,
              mean=None, std=None) -> Optional[torch.Tensor]:
    """In-place fused PGD step on ``image`` (fp32). Returns the normalised image when mean/std given."""
    lib = load()
    for n, t in (("image", image), ("grad", grad), ("orig", orig)):
        require_cuda(t, n)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"bimodal_b200.pgd_step_: `{n}` must be contiguous float32")
    if not (image.shape == grad.shape == orig.shape):
        raise RuntimeError("bimodal_b200.pgd_step_: shape mismatch")
    norm_out = None
    mean_arr = std_arr = None
    channels, plane = 0, 1
    if mean is not None:
        channels = image.shape[-3]
        plane = image.shape[-1] * image.shape[-2]
        mean_l = [float(torch.tensor(m, dtype=torch.float32)) for m in mean]
        std_l = [float(torch.tensor(s, dtype=torch.float32)) for s in std]
        if len(mean_l) != channels or len(std_l) != channels or channels > 4:
            raise RuntimeError("bimodal_b200.pgd_step_: mean/std must have one entry per channel (<=4)")
        mean_arr = (C.c_float * channels)(*mean_l)
        std_arr = (C.c_float * channels)(*std_l)
        norm_out = torch.empty_like(image)
    with torch.cuda.device(image.device):
        check(lib.bma_pgd_step(image.data_ptr(), grad.data_ptr(), orig.data_ptr(), image.numel(),
                               float(step), float(eps), ptr(norm_out), mean_arr, std_arr, channels,
                               plane, stream_ptr(image.device)), "pgd_step")
    return norm_out


def token_grad(grad_embeds: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``grad_embeds [n_tok, d] @ weight[V, d].T`` -> ``[n_tok, V]`` bf16 on the tcgen05 GEMM."""
    lib = load()
    require_cuda(grad_embeds, "grad_embeds")
    require_cuda(weight, "weight")
    if grad_embeds.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise RuntimeError("bimodal_b200.token_grad: bf16 operands required")
    g = grad_embeds.contiguous()
    n_tok, d = g.shape
    V = weight.shape[0]
    if weight.stride(1) != 1 or weight.shape[1] != d:
        raise RuntimeError("bimodal_b200.token_grad: weight must be [V, d] with unit inner stride")
    out = torch.empty((n_tok, V), dtype=torch.bfloat16, device=g.device)
    with torch.cuda.device(g.device):
        check(lib.bma_token_grad(g.data_ptr(), g.stride(0), weight.data_ptr(), weight.stride(0),
                                 out.data_ptr(), out.stride(0), n_tok, d, V, stream_ptr(g.device)),
              "token_grad")
    return out


def pack_mask_bits(not_allowed_ids: Optional[torch.Tensor], vocab: int, device) -> Optional[torch.Tensor]:
    """V-bit set (int32 words, little-endian bit order) of the not-allowed token ids."""
    if not_allowed_ids is None:
        return None
    words = (vocab + 31) // 32
    flags = torch.zeros(words * 32, dtype=torch.bool, device=device)
    ids = not_allowed_ids.to(device=device, dtype=torch.long)
    flags[ids] = True
    weights = (1 << torch.arange(32, device=device, dtype=torch.int64))
    packed = (flags.view(words, 32).to(torch.int64) * weights).sum(dim=1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)
    return packed.to(torch.int32).contiguous()


def sample_ids(ids: torch.Tensor, grad: torch.Tensor, search_width: int, topk: int, n_replace: int,
               mask_bits: Optional[torch.Tensor], seed: int, offset: int,
               want_topk: bool = False):
    """Device sampler. Returns (new_ids [sw, n_tok] int64, offset_advance, topk_ids or None)."""
    lib = load()
    require_cuda(grad, "grad")
    require_cuda(ids, "ids")
    if grad.dim() != 2 or grad.stride(1) != 1:
        raise RuntimeError("bimodal_b200.sample_ids: grad must be [n_tok, V] with unit inner stride")
    if grad.dtype not in DTYPE_CODES:
        raise RuntimeError(f"bimodal_b200.sample_ids: unsupported grad dtype {grad.dtype}")
    n_tok, V = grad.shape
    if topk > V:
        raise RuntimeError("selected index k out of range")   # what torch.topk raises
    ids = ids.to(torch.int64).contiguous()
    dev = grad.device
    ws = torch.empty(lib.bma_sampler_workspace_bytes(n_tok, V, topk), dtype=torch.uint8, device=dev)
    new_ids = torch.empty((search_width, n_tok), dtype=torch.int64, device=dev)
    topk_ids = torch.empty((n_tok, topk), dtype=torch.int64, device=dev) if want_topk else None
    adv = _ULL(0)
    with torch.cuda.device(dev):
        check(lib.bma_sample_ids(ids.data_ptr(), grad.data_ptr(), DTYPE_CODES[grad.dtype],
                                 grad.stride(0), n_tok, V, ptr(mask_bits), search_width, topk,
                                 n_replace, seed, offset, ws.data_ptr(), ptr(topk_ids),
                                 new_ids.data_ptr(), C.byref(adv), stream_ptr(dev)), "sample_ids")
    return new_ids, int(adv.value), topk_ids
