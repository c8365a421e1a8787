"""ctypes binding of librpo_b200.so (include/rpo_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this module
raises.  `load()` never builds implicitly on a machine without nvcc; run `python -m rpo_b200.build`
(or `__graft_entry__.build()`) first.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "librpo_b200.so")

RPO_F32, RPO_F16, RPO_BF16 = 0, 1, 2
RPO_U8 = 3  # image dtype of rpo_forward only
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2
ACT_NONE, ACT_QUICKGELU = 0, 1

# every symbol include/rpo_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "rpo_last_error", "rpo_version", "rpo_create", "rpo_destroy", "rpo_device_bytes", "rpo_bind_weights",
    "rpo_set_classes", "rpo_set_image_norm", "rpo_forward", "rpo_backward", "rpo_sgd_step", "rpo_layernorm_fwd", "rpo_layernorm_bwd",
    "rpo_gemm_bias_act", "rpo_ro_attention_fwd", "rpo_ro_attention_fwd_dense", "rpo_ro_attention_fwd_dense_supported", "rpo_ro_attention_bwd", "rpo_logits_ce_fwd", "rpo_logits_ce_bwd",
    "rpo_debug_fetch", "rpo_launch_count", "rpo_profile_begin", "rpo_profile_end",
    "rpo_bind_text_exchange", "rpo_forward_text", "rpo_forward_image", "rpo_forward_logits", "rpo_backward_logits",
    "rpo_backward_text", "rpo_backward_image", "rpo_forward_image_context", "rpo_forward_image_prompts",
    "rpo_peer_signal_bytes", "rpo_peer_epoch_bytes", "rpo_peer_allreduce_sgd", "rpo_peer_all_gather",
    "rpo_peer_reduce_scatter",
]


class RpoConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "dtype", "K", "n_cls", "ctx_len", "embed_dim", "v_width", "v_layers", "v_heads", "v_patch", "v_res",
        "t_width", "t_layers", "t_heads", "max_batch", "gemm_backend", "cls_first", "cls_local", "image_slots")]


class RpoBlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "in_w", "in_b", "out_w", "out_b", "fc_w", "fc_b", "proj_w", "proj_b")]


class RpoWeights(C.Structure):
    _fields_ = [("v_blocks", C.POINTER(RpoBlockWeights)), ("t_blocks", C.POINTER(RpoBlockWeights))] + \
        [(n, C.c_void_p) for n in (
            "conv_w", "cls_emb", "v_pos", "ln_pre_w", "ln_pre_b", "ln_post_w", "ln_post_b", "v_proj",
            "ln_final_w", "ln_final_b", "t_proj", "logit_scale")]


PEER_MAX_WORLD = 8


class RpoPeerComm(C.Structure):
    _fields_ = [("signals", C.c_void_p * PEER_MAX_WORLD), ("epoch", C.c_void_p), ("rank", C.c_int32),
                ("world", C.c_int32)]


class RpoError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the shared library once.  Raises if it is absent -- the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RpoError(f"{LIB_PATH} not found: build it with `python -m rpo_b200.build` "
                       f"(nvcc, sm_100a); rpo_b200 has no CPU/PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.rpo_last_error.restype = C.c_char_p
    lib.rpo_version.restype = C.c_int
    lib.rpo_create.argtypes = [C.POINTER(RpoConfig), C.POINTER(vp)]
    lib.rpo_destroy.argtypes = [vp]
    lib.rpo_destroy.restype = None
    lib.rpo_device_bytes.argtypes = [vp]
    lib.rpo_device_bytes.restype = C.c_size_t
    lib.rpo_bind_weights.argtypes = [vp, C.POINTER(RpoWeights), vp]
    lib.rpo_set_classes.argtypes = [vp, vp, C.POINTER(i32), vp]
    lib.rpo_forward.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.rpo_set_image_norm.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.rpo_backward.argtypes = [vp, vp, vp]
    lib.rpo_bind_text_exchange.argtypes = [vp, vp, vp]
    lib.rpo_forward_text.argtypes = [vp, vp, vp]
    lib.rpo_forward_image.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.rpo_forward_logits.argtypes = [vp, vp, vp, vp, vp]
    lib.rpo_forward_image_context.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.rpo_forward_image_prompts.argtypes = [vp, vp, i32, vp]
    lib.rpo_peer_signal_bytes.restype = C.c_size_t
    lib.rpo_peer_epoch_bytes.restype = C.c_size_t
    lib.rpo_peer_allreduce_sgd.argtypes = [C.POINTER(RpoPeerComm), C.POINTER(vp), vp, vp, i32, i64, i64, vp, vp, f32,
                                           f32, f32, vp, vp]
    lib.rpo_peer_all_gather.argtypes = [C.POINTER(RpoPeerComm), C.POINTER(vp), i64, i64, i64, vp]
    lib.rpo_peer_reduce_scatter.argtypes = [C.POINTER(RpoPeerComm), C.POINTER(vp), i32, i64, i64, i64, vp]
    lib.rpo_backward_logits.argtypes = [vp, vp]
    lib.rpo_backward_text.argtypes = [vp, vp, vp]
    lib.rpo_backward_image.argtypes = [vp, vp, vp]
    lib.rpo_sgd_step.argtypes = [vp, i32, vp, vp, i64, vp, f32, f32, f32, vp, vp]
    lib.rpo_layernorm_fwd.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp]
    lib.rpo_layernorm_bwd.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp]
    lib.rpo_gemm_bias_act.argtypes = [vp, i64, vp, i64, vp, i64, i64, i32, i32, vp, i32, vp, vp, vp, i64, i32, i32, vp]
    lib.rpo_ro_attention_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.rpo_ro_attention_fwd_dense.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.rpo_ro_attention_fwd_dense_supported.argtypes = [i32, i32, i32, i32]
    lib.rpo_ro_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.rpo_logits_ce_fwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.rpo_logits_ce_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp,
                                      i32, vp]
    lib.rpo_debug_fetch.argtypes = [vp, i32, i32, vp, i64, vp]
    lib.rpo_debug_fetch.restype = i64
    lib.rpo_launch_count.argtypes = [vp]
    lib.rpo_launch_count.restype = i64
    lib.rpo_profile_begin.argtypes = [vp]
    lib.rpo_profile_end.argtypes = [C.c_char_p, i64]
    lib.rpo_profile_end.restype = i64
    for name in SYMBOLS:
        getattr(lib, name)  # raises if the .so lacks a symbol the header declares
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        msg = load().rpo_last_error()
        raise RpoError(f"librpo_b200 call failed (status {status}): {msg.decode() if msg else '?'}")


def dtype_code(torch_dtype):
    import torch
    return {torch.float32: RPO_F32, torch.float16: RPO_F16, torch.bfloat16: RPO_BF16}[torch_dtype]


def ptr(t):
    """device pointer of a (contiguous) torch tensor, or NULL"""
    if t is None:
        return None
    assert t.is_contiguous(), "librpo_b200 takes contiguous buffers"
    return t.data_ptr()


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
