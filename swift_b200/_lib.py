"""ctypes binding of ``libswift_b200.so`` (the C ABI declared in ``include/swift_b200.h``).

There is deliberately no fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SWB_LIB", os.path.join(HERE, "libswift_b200.so"))   # SWB_LIB: A/B builds (tools only)

ABI_VERSION = 14

# Every symbol ``include/swift_b200.h`` declares; tests check the library exports exactly these.
EXPORTS = (
    "swb200_abi_version", "swb200_last_error", "swb200_validate", "swb200_workspace_bytes",
    "swb200_conditioning_scratch_bytes", "swb200_conditioning", "swb200_forward", "swb200_gemm",
    "swb200_gemm_qkv", "swb200_gemm_swiglu", "swb200_gemm_embed", "swb200_gemm_head", "swb200_patch_gather",
    "swb200_ln_mod_residual", "swb200_window_attention", "swb200_rollout_noise", "swb200_rollout_forcings",
    "swb200_rollout_advance", "swb200_trace_enable", "swb200_trace_report", "swb200_ln_workspace_bytes",
    "swb200_gemm_ln_residual", "swb200_ensemble_stats", "swb200_jvp_workspace_bytes",
    "swb200_conditioning_jvp_scratch_bytes", "swb200_conditioning_jvp", "swb200_forward_jvp",
    "swb200_scm_noised_inputs", "swb200_scm_target_scratch_bytes", "swb200_scm_tangent_target",
    "swb200_debug_saturation",
    "swb200_train_tape_bytes", "swb200_train_workspace_bytes", "swb200_train_forward", "swb200_train_backward_head",
    "swb200_train_backward_layer", "swb200_train_backward_embed", "swb200_conditioning_backward_scratch_bytes",
    "swb200_conditioning_backward", "swb200_gemm_splitk", "swb200_transpose16", "swb200_ln_backward_scratch_bytes",
    "swb200_ln_backward", "swb200_swiglu_backward", "swb200_attention_backward_scratch_bytes", "swb200_attention_backward",
    "swb200_qkv_pack_train", "swb200_muon_workspace_bytes", "swb200_muon_step", "swb200_adam_step", "swb200_muon_vector_step",
    "swb200_packed_bytes", "swb200_pack_weights", "swb200_train_packed_bytes", "swb200_pack_train_weights",
    "swb200_conditioning_backward_logvar", "swb200_logvar_head", "swb200_scm_tangent_target_logvar",
    "swb200_scm_distill_direction",
)

_i32, _f32, _vp, _sz = C.c_int32, C.c_float, C.c_void_p, C.c_size_t


class Model(C.Structure):
    """``struct swb200_model`` (field order must match include/swift_b200.h)."""
    _fields_ = (
        [(n, _i32) for n in ("img_h", "img_w", "patch_h", "patch_w", "win_h", "win_w", "shift_h", "shift_w",
                             "in_channels", "out_channels", "depth", "dim", "heads", "dff", "aux_dim",
                             "k_embed", "split_embed", "split_head", "gemm_tile", "attn_impl", "act_fp16",
                             "fuse_ln", "attn_fp16", "x_single")]
        + [("timestep_weight", _f32)]
        + [(n, _vp) for n in ("w_embed", "b_embed", "pos_embed", "aux_w", "aux_b", "l1_w", "l1_b", "l2_w", "l2_b",
                              "mod_w", "mod_b", "ln_gamma", "ln_beta", "qscale", "w_qkv", "w_o", "w_1", "w_2",
                              "w_head")]
    )


class Update(C.Structure):
    """``struct swb200_update``: y = alpha*xt + beta*F + gamma*fprev; optional raw F output."""
    _fields_ = [("xt", _vp), ("fprev", _vp), ("out_f", _vp), ("alpha", _f32), ("beta", _f32), ("gamma", _f32),
                ("state", _vp), ("state_channels", _i32), ("zero_channel", _i32), ("x_std", _vp), ("x_mean", _vp),
                ("d_std", _vp), ("phys", _vp)]


class RefParams(C.Structure):
    """``struct swb200_ref_params``: device pointers to a reference checkpoint's fp32 parameters."""
    _PP = C.POINTER(C.c_void_p)
    _fields_ = ([(n, _vp) for n in ("pos_embed", "patch_w", "patch_b", "aux_w", "aux_b", "l1_w", "l1_b", "l2_w", "l2_b", "head_w")]
                + [(n, C.POINTER(C.c_void_p)) for n in ("scale", "attn_ln_w", "attn_ln_b", "attn_mod_w", "attn_mod_b", "to_qkv", "wo",
                                                         "ff_ln_w", "ff_ln_b", "ff_mod_w", "ff_mod_b", "w1", "w2")])


class TrainModel(C.Structure):
    """``struct swb200_train_model``: the bf16 base model + plain / transposed bf16 copies of the per-layer matrices."""
    _fields_ = ([("base", Model)]
                + [(n, _vp) for n in ("w_qkv", "w_o", "w_1", "w_2", "wt_qkv", "wt_o", "wt_1", "wt_2", "wt_head")]
                + [("kp_head", _i32)])


class TrainGrads(C.Structure):
    """``struct swb200_train_grads`` (fp32 device buffers)."""
    _fields_ = ([(n, _vp) for n in ("w_qkv", "w_o", "w_1", "w_2", "w_head", "w_embed_t", "b_embed", "pos_embed", "dscale",
                                    "dgain", "dbias")] + [("accumulate", _i32)])


class CondGrads(C.Structure):
    """``struct swb200_cond_grads``."""
    _fields_ = [(n, _vp) for n in ("aux_w", "aux_b", "l1_w", "l1_b", "l2_w", "l2_b", "mod_w", "mod_b", "ln_gamma", "ln_beta")]


_ANY_ABI = bool(os.environ.get("SWB_LIB_ANYABI"))     # tools only: time an older build of the library (A/B of kernels)
_lock = threading.Lock()
_lib = None


def _declare(lib):
    MP, UP = C.POINTER(Model), C.POINTER(Update)
    TP, GP, CP = C.POINTER(TrainModel), C.POINTER(TrainGrads), C.POINTER(CondGrads)
    sig = {
        "swb200_packed_bytes": (_sz, [MP]),
        "swb200_pack_weights": (C.c_int, [MP, C.POINTER(RefParams), _vp, _sz, _vp]),
        "swb200_train_packed_bytes": (_sz, [TP]),
        "swb200_pack_train_weights": (C.c_int, [TP, C.POINTER(RefParams), _vp, _sz, _vp]),
        "swb200_train_tape_bytes": (_sz, [TP, C.c_int]),
        "swb200_train_workspace_bytes": (_sz, [TP, C.c_int]),
        "swb200_train_forward": (C.c_int, [TP, _vp, C.c_int, _f32, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
        "swb200_train_backward_head": (C.c_int, [TP, C.c_int, _vp, _vp, _vp, _sz, GP, _vp]),
        "swb200_train_backward_layer": (C.c_int, [TP, C.c_int, C.c_int, _vp, _vp, _vp, _sz, GP, _vp]),
        "swb200_train_backward_embed": (C.c_int, [TP, C.c_int, _vp, _vp, _sz, GP, _vp]),
        "swb200_conditioning_backward_scratch_bytes": (_sz, [MP, C.c_int]),
        "swb200_conditioning_backward": (C.c_int, [MP, _vp, C.c_int, _vp, _vp, _vp, CP, C.c_int, _vp, _sz, _vp]),
        "swb200_conditioning_backward_logvar": (C.c_int, [MP, _vp, C.c_int, _vp, _vp, _vp, CP, C.c_int, _vp, _sz, _vp, _vp, _vp, _vp,
                                                          _vp]),
        "swb200_scm_distill_direction": (C.c_int, [_vp, _vp, _f32, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
        "swb200_logvar_head": (C.c_int, [MP, _vp, _vp, _vp, C.c_int, _vp, _vp]),
        "swb200_scm_tangent_target_logvar": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                                       C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
        "swb200_gemm_splitk": (C.c_int, [C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         _vp]),
        "swb200_transpose16": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp, C.c_int64, _vp]),
        "swb200_ln_backward_scratch_bytes": (_sz, [C.c_int, C.c_int, C.c_int]),
        "swb200_ln_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _sz, _vp]),
        "swb200_swiglu_backward": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
        "swb200_attention_backward_scratch_bytes": (_sz, [C.c_int, C.c_int, C.c_int, C.c_int]),
        "swb200_attention_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_int, _vp, _sz, _vp]),
        "swb200_qkv_pack_train": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
        "swb200_muon_workspace_bytes": (_sz, [C.c_int, C.c_int, C.c_int]),
        "swb200_muon_step": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int,
                                       _f32, _f32, _f32, C.c_int, C.c_int, _vp, _sz, _vp]),
        "swb200_muon_vector_step": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _f32, _f32, _f32, C.c_int, C.c_int, _vp]),
        "swb200_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, _f32, _f32, _f32, _f32, _f32, C.c_int, _vp]),
        "swb200_abi_version": (C.c_int, []),
        "swb200_last_error": (C.c_char_p, []),
        "swb200_validate": (C.c_int, [MP]),
        "swb200_workspace_bytes": (_sz, [MP, C.c_int]),
        "swb200_conditioning_scratch_bytes": (_sz, [MP, C.c_int]),
        "swb200_conditioning": (C.c_int, [MP, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
        "swb200_forward": (C.c_int, [MP, _vp, C.c_int, _f32, _vp, C.c_int, C.c_int, _vp, _vp, UP, _vp, _vp, _sz, _vp]),
        "swb200_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                  C.c_int, C.c_int, _vp]),
        "swb200_gemm_qkv": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
        "swb200_gemm_swiglu": (C.c_int, [C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
        "swb200_gemm_embed": (C.c_int, [C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _vp,
                                        C.c_int, C.c_int, _vp]),
        "swb200_gemm_head": (C.c_int, [C.c_int, MP, _vp, C.c_int, C.c_int, C.c_int, UP, _vp, _vp]),
        "swb200_patch_gather": (C.c_int, [MP, _vp, C.c_int, _f32, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]),
        "swb200_ln_mod_residual": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
        "swb200_window_attention": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, _vp, _vp]),
        "swb200_rollout_noise": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int64, _vp]),
        "swb200_rollout_forcings": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int,
                                              C.c_int, _vp]),
        "swb200_rollout_advance": (C.c_int, [_vp, _vp]),
        "swb200_ln_workspace_bytes": (_sz, [C.c_int, C.c_int]),
        "swb200_gemm_ln_residual": (C.c_int, [C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int,
                                              C.c_int, _vp, C.c_int, _vp]),
        "swb200_ensemble_stats": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int,
                                            _vp, _vp]),
        "swb200_jvp_workspace_bytes": (_sz, [MP]),
        "swb200_conditioning_jvp_scratch_bytes": (_sz, [MP, C.c_int]),
        "swb200_conditioning_jvp": (C.c_int, [MP, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
        "swb200_forward_jvp": (C.c_int, [MP, _vp, C.c_int, _f32, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _sz, _vp]),
        "swb200_scm_noised_inputs": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
        "swb200_scm_target_scratch_bytes": (_sz, [C.c_int]),
        "swb200_scm_tangent_target": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, C.c_int, C.c_int, C.c_int,
                                                C.c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
        "swb200_debug_saturation": (C.c_int, [_vp]),
        "swb200_trace_enable": (C.c_int, [C.c_int]),
        "swb200_trace_report": (C.c_int, [C.c_char_p, _sz]),
    }
    assert set(sig) == set(EXPORTS)
    for name, (res, args) in sig.items():
        if _ANY_ABI and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    """Load (once) and return the CUDA library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the swift_b200 CUDA library has not been built "
                        "(run `python -m swift_b200.build`); there is no CPU or PyTorch fallback")
                l = C.CDLL(LIB_PATH)
                _declare(l)
                if l.swb200_abi_version() != ABI_VERSION and not _ANY_ABI:
                    raise RuntimeError("libswift_b200.so ABI version mismatch; rebuild with `python -m swift_b200.build`")
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().swb200_last_error()
        raise RuntimeError(f"swift_b200 {what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
