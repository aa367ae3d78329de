"""ctypes binding of include/msf_b200.h (the C ABI).  No torch types cross it:
only raw device pointers, sizes and the CUDA stream handle."""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t,
                    c_uint64, c_void_p)

MSF_MAX_MODALITIES = 8
MSF_PREC_F32, MSF_PREC_BF16 = 0, 1
MSF_TRAIN_DEAD_SLOTS_ZERO = 1
MSF_OPT_NORM_GIVEN = 2
MSF_ABI_VERSION = 4
MSF_FOLD_REUSE_PROJECTIONS = 1
MSF_FOLD_PROJECTIONS_ONLY = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MsfError(RuntimeError):
    """A C-ABI call returned a negative status (message from msf_last_error())."""


class FusionShape(Structure):
    _fields_ = [
        ("num_modalities", c_int32),
        ("hidden", c_int32),
        ("num_heads", c_int32),
        ("num_classes", c_int32),
        ("in_dims", c_int32 * MSF_MAX_MODALITIES),
        ("pair_present", c_uint64),
    ]


class FusionCall(Structure):
    _fields_ = [
        ("batch", c_int32),
        ("precision", c_int32),
        ("training", c_int32),
        ("dropout_p", c_float),
        ("seed", c_uint64),
        ("offset", c_uint64),
        ("rng_state", c_void_p),
        ("params", c_void_p),
        ("params_bf16", c_void_p),
        ("x", c_void_p * MSF_MAX_MODALITIES),
        ("mask", c_void_p),
        ("workspace", c_void_p),
        ("workspace_bytes", c_size_t),
        ("logits", c_void_p),
        ("fusion_weights", c_void_p),
        ("attn_gates", c_void_p),
        ("grad_logits", c_void_p),
        ("grad_params", c_void_p),
        ("grad_x", c_void_p * MSF_MAX_MODALITIES),
        ("grad_sq", c_void_p),
        ("ln_weight", c_void_p * MSF_MAX_MODALITIES),
        ("ln_bias", c_void_p * MSF_MAX_MODALITIES),
        ("ln_eps", c_float),
        ("x_bf16", c_int32),
    ]


class DpComm(Structure):
    _fields_ = [
        ("rank", c_int32),
        ("world", c_int32),
        ("grads", c_void_p * 8),
        ("stages", c_void_p * 8),
        ("reds", c_void_p * 8),
        ("sigs", c_void_p * 8),
    ]


class DpzComm(Structure):  # msf_dpz_comm
    _fields_ = [
        ("rank", c_int32),
        ("world", c_int32),
        ("stages", c_void_p * 8),
        ("arenas_bf16", c_void_p * 8),
        ("params", c_void_p * 8),
        ("sigs", c_void_p * 8),
        ("mc_grad", c_void_p),
        ("mc_arena_bf16", c_void_p),
        ("mc_params", c_void_p),
    ]


class LstmSeq(ctypes.Structure):  # msf_lstm_seq
    _fields_ = [
        ("x_bf16", c_void_p), ("w_hh", c_void_p), ("w_ih", c_void_p), ("bias", c_void_p),
        ("h_a", c_void_p), ("h_b", c_void_p), ("cell", c_void_p), ("h_out", c_void_p), ("lengths", c_void_p),
        # training mode
        ("h_all", c_void_p), ("z_in", c_void_p), ("gates", c_void_p), ("c_all", c_void_p), ("w_hh_t", c_void_p), ("d_h_out", c_void_p), ("d_h_all", c_void_p),
        ("dc", c_void_p), ("partial", c_void_p), ("d_w_ih", c_void_p), ("d_w_hh", c_void_p), ("d_bias", c_void_p),
        ("features", c_int32), ("in_cols", c_int32), ("cell_type", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/msf_b200.h declares
PROTOTYPES = {
    "msf_abi_version": (c_int32, []),
    "msf_dpz_optimizer_step_packed": (c_int32, [POINTER(FusionShape), POINTER(DpzComm), c_void_p, c_void_p, c_void_p,
                                                c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                                                c_float, c_int32, c_void_p]),
    "msf_dpz_owner_map": (c_int32, [POINTER(FusionShape), c_int32, c_void_p]),
    "msf_fusion_layer_norm_fused": (c_int32, [POINTER(FusionShape), c_int32]),
    "msf_layer_norm_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p]),
    "msf_layer_norm_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                          c_float, c_void_p]),
    "msf_struct_sizes": (c_int32, [POINTER(c_int32), POINTER(c_int32)]),
    "msf_last_error": (c_char_p, []),
    "msf_launch_count": (c_uint64, []),
    "msf_prof_enable": (c_int32, [c_int32]),
    "msf_prof_report": (c_int32, [ctypes.c_char_p, c_size_t]),
    "msf_memcpy_batch": (c_int32, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_size_t), c_int32, c_void_p]),
    "msf_device_check": (c_int32, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "msf_fusion_param_count": (c_int32, [POINTER(FusionShape), POINTER(c_int64)]),
    "msf_fusion_param_offset": (c_int32, [POINTER(FusionShape), c_int32, c_int32, POINTER(c_int64)]),
    "msf_fusion_workspace_bytes": (c_int32, [POINTER(FusionShape), c_int32, c_int32, POINTER(c_size_t)]),
    "msf_fusion_bf16_arena_bytes": (c_int32, [POINTER(FusionShape), POINTER(c_size_t)]),
    "msf_fusion_pack_bf16": (c_int32, [POINTER(FusionShape), c_void_p, c_void_p, c_void_p]),
    "msf_fusion_forward": (c_int32, [POINTER(FusionShape), POINTER(FusionCall), c_void_p]),
    "msf_fusion_backward": (c_int32, [POINTER(FusionShape), POINTER(FusionCall), c_void_p]),
    "msf_fusion_train_pass": (c_int32, [POINTER(FusionShape), POINTER(FusionCall), c_void_p, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "msf_fusion_train_pass_is_fused": (c_int32, [POINTER(FusionShape), c_int32]),
    "msf_fusion_infer_pass": (c_int32, [POINTER(FusionShape), POINTER(FusionCall), c_void_p, c_void_p, ctypes.c_uint32,
                                        c_void_p]),
    "msf_debug_head_stamps": (c_int32, [c_void_p]),
    "msf_debug_chain_stamps": (c_int32, [c_void_p]),
    "msf_debug_proj_stamps": (c_int32, [c_void_p]),
    "msf_adaptive_weights": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                       c_void_p, c_void_p]),
    "msf_dropout_mask": (c_int32, [c_uint64, c_uint64, c_int32, c_int32, c_int64, c_int64, c_float,
                                   c_void_p, c_void_p]),
    "msf_arena_gather": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "msf_arena_scatter": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "msf_cross_entropy": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_float, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "msf_softmax_conf_pred": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "msf_ece_bin": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double), c_int32,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_grad_sq_norm": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "msf_adamw_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float,
                                 c_float, c_float, c_float, c_float, c_float, c_float, c_void_p,
                                 c_void_p]),
    "msf_adamw_step_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float,
                                     c_float, c_float, c_float, c_float, c_float, c_float, c_void_p,
                                     c_void_p]),
    "msf_fusion_optimizer_step": (c_int32, [POINTER(FusionShape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_float, c_float, c_float, c_float, c_float, c_float, c_float,
                                            c_void_p, c_void_p]),
    "msf_fusion_optimizer_step_packed": (c_int32, [POINTER(FusionShape), c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                                                   c_float, c_void_p, c_void_p, c_int32, c_void_p]),
    "msf_dp_optimizer_step": (c_int32, [POINTER(FusionShape), POINTER(DpComm), c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_float, c_float, c_float, c_float, c_float, c_float, c_float, c_void_p]),
    "msf_dp_optimizer_step_packed": (c_int32, [POINTER(FusionShape), POINTER(DpComm), c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                                               c_float, c_void_p, c_int32, c_void_p]),
    "msf_train_state_advance": (c_int32, [c_void_p, c_void_p]),
    "msf_linear_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                     c_int32, c_void_p]),
    "msf_linear_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "msf_attention_core_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                             c_int32, c_int32, c_int32, c_float, c_int32, c_uint64,
                                             c_uint64, c_void_p, c_void_p, c_void_p]),
    "msf_attention_core_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                              c_int32, c_int32, c_int32, c_float, c_int32, c_uint64,
                                              c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p]),
    "msf_lstm_forward": (c_int32, [POINTER(LstmSeq), c_int32, c_int64, c_int32, c_int32, c_void_p]),
    "msf_lstm_seq_bytes": (c_int32, []),
    "msf_lstm_pack_input": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "msf_lstm_backward": (c_int32, [POINTER(LstmSeq), c_int32, c_int64, c_int32, c_int32, c_void_p]),
    "msf_lstm_f32_forward": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                       c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_lstm_f32_backward": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "msf_gru_f32_forward": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                      c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_gru_f32_backward": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "msf_frame_pool_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "msf_frame_pool_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_lstm_dropout": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_uint64, c_uint64, c_int32, c_void_p]),
    "msf_lstm_backward_scratch_bytes": (c_int32, [c_int64, c_int32, c_int32, POINTER(ctypes.c_size_t)]),
    "msf_fusion_infer_folded": (c_int32, [POINTER(FusionShape), POINTER(FusionCall), c_void_p, c_void_p, ctypes.c_uint32,
                                          c_int32, c_void_p, c_void_p, c_void_p]),
    "msf_eval_accumulate": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, POINTER(ctypes.c_double), c_int32, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_bad_label_count": (c_int32, [POINTER(c_int64), c_int32]),
    "msf_grad_accumulate": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_int32, c_void_p]),
    "msf_bn_act_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                     c_float, c_int32, c_int32, c_float, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_bn_act_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int32,
                                      c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "msf_gemm_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64,
                                c_int64, c_int64, c_int64, c_int32, c_void_p, c_int32, c_void_p]),
}


def library_path() -> str:
    return os.environ.get("MSF_B200_LIB", os.path.join(_HERE, "libmsf_b200.so"))


def lib():
    """Load libmsf_b200.so (once).  Fails loudly: there is no fallback path."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise MsfError(
                f"{path} not found: build it with `python {os.path.join(_HERE, 'build.py')}` "
                "(nvcc, sm_100a). There is no CPU / PyTorch fallback for this path."
            )
        handle = ctypes.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.msf_abi_version() != MSF_ABI_VERSION:
            raise MsfError(f"ABI mismatch: library {handle.msf_abi_version()}, binding {MSF_ABI_VERSION}")
        a, b = c_int32(), c_int32()
        handle.msf_struct_sizes(ctypes.byref(a), ctypes.byref(b))
        if (a.value, b.value) != (ctypes.sizeof(FusionShape), ctypes.sizeof(FusionCall)):
            raise MsfError(f"struct layout mismatch: library {(a.value, b.value)}, binding "
                           f"{(ctypes.sizeof(FusionShape), ctypes.sizeof(FusionCall))}")
        if handle.msf_lstm_seq_bytes() != ctypes.sizeof(LstmSeq):
            raise MsfError(f"struct layout mismatch: msf_lstm_seq is {handle.msf_lstm_seq_bytes()} bytes in the library, "
                           f"{ctypes.sizeof(LstmSeq)} in the binding")
        _LIB = handle
    return _LIB


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().msf_last_error()
        raise MsfError(f"msf_b200 error {rc}: {msg.decode() if msg else '?'}")
