"""Frozen CLIP weights for the RPO hot path: precision rule, a minimal CLIP stand-in built from a
state dict (tests / bench have no real checkpoint), and packing into two flat device buffers whose
addresses are handed to librpo_b200 (RpoWeights in include/rpo_b200.h).
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

PREC_DTYPE = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}

_BLOCK_16 = ("attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
             "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")
_TOP_16 = ("visual.conv1.weight", "text_projection", "visual.proj")


def is_matmul_weight(key: str) -> bool:
    """Tensors the reference stores in the model dtype: Linear / Conv / MultiheadAttention weights and
    biases and the two projection matrices (clip/model.py:379-400).  Everything else (LayerNorm
    affine, embeddings, class embedding, logit_scale) stays fp32 and is cast at use."""
    return key.endswith(_BLOCK_16) or key in _TOP_16


def convert_state_dict(sd, prec: str):
    """fp32 master state dict -> what `clip.model.build_model` (+ `.float()` for PREC=fp32,
    trainers/rpo.py:247-249) leaves in memory: matmul weights rounded through fp16 (held as fp32 when
    prec == 'fp32'), or in bf16 for the bf16 extension."""
    dt = PREC_DTYPE[prec]
    out = {}
    for k, v in sd.items():
        if is_matmul_weight(k):
            out[k] = v.to(torch.float16).float() if dt == torch.float32 else v.to(dt)
        else:
            out[k] = v.float() if v.is_floating_point() else v
    return out


class SyntheticCLIP(nn.Module):
    """Just enough of `clip.model.CLIP` for CustomCLIP / PromptLearner: `state_dict()` with the CLIP key
    set, `.dtype` (clip/model.py:340-342) and `.visual.input_resolution`."""

    def __init__(self, state_dict, prec: str = "fp16"):
        super().__init__()
        self._sd = convert_state_dict(state_dict, prec)
        p = self._sd["visual.conv1.weight"].shape[-1]
        grid = round((self._sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        self.visual = SimpleNamespace(input_resolution=p * grid)

    @property
    def dtype(self):
        return self._sd["visual.conv1.weight"].dtype

    def state_dict(self, *args, **kwargs):
        return self._sd


def arch_from_state_dict(sd):
    """Same inference as clip/model.py:403-440 build_model, ViT branch only."""
    if "visual.proj" not in sd:
        raise ValueError("RPO needs a ViT visual backbone (trainers/rpo.py:52,154 assume ViT token shapes)")
    v_width = sd["visual.conv1.weight"].shape[0]
    v_layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    patch = sd["visual.conv1.weight"].shape[-1]
    grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    t_width = sd["ln_final.weight"].shape[0]
    t_layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")})
    return SimpleNamespace(
        embed_dim=sd["text_projection"].shape[1], v_width=v_width, v_layers=v_layers, v_heads=v_width // 64,
        v_patch=patch, v_res=patch * grid, ctx_len=sd["positional_embedding"].shape[0], t_width=t_width,
        t_layers=t_layers, t_heads=t_width // 64)


_BLOCK_FIELDS = (  # (RpoBlockWeights field, state-dict suffix, is fp32)
    ("ln1_w", "ln_1.weight", True), ("ln1_b", "ln_1.bias", True), ("ln2_w", "ln_2.weight", True),
    ("ln2_b", "ln_2.bias", True), ("in_w", "attn.in_proj_weight", False), ("in_b", "attn.in_proj_bias", False),
    ("out_w", "attn.out_proj.weight", False), ("out_b", "attn.out_proj.bias", False),
    ("fc_w", "mlp.c_fc.weight", False), ("fc_b", "mlp.c_fc.bias", False), ("proj_w", "mlp.c_proj.weight", False),
    ("proj_b", "mlp.c_proj.bias", False))
_TOP_FIELDS = (
    ("conv_w", "visual.conv1.weight", False), ("cls_emb", "visual.class_embedding", True),
    ("v_pos", "visual.positional_embedding", True), ("ln_pre_w", "visual.ln_pre.weight", True),
    ("ln_pre_b", "visual.ln_pre.bias", True), ("ln_post_w", "visual.ln_post.weight", True),
    ("ln_post_b", "visual.ln_post.bias", True), ("v_proj", "visual.proj", False),
    ("ln_final_w", "ln_final.weight", True), ("ln_final_b", "ln_final.bias", True),
    ("t_proj", "text_projection", False), ("logit_scale", "logit_scale", True))

_ALIGN = 128  # elements; keeps every tensor 256/512-byte aligned (TMA needs 16 B)


def pack_weights(sd, arch, dtype):
    """Packs the frozen weights used by the path into one `dtype` buffer and one fp32 buffer.
    Returns (w_mm, w_f32, index) where index maps state-dict key -> (which, offset, numel)."""
    mm, f32, index = [], [], {}
    cur = {"mm": 0, "f32": 0}

    def add(key, which):
        t = sd[key].detach().reshape(-1)
        t = t.to(dtype) if which == "mm" else t.float()
        pad = (-t.numel()) % _ALIGN
        (mm if which == "mm" else f32).append(torch.cat([t.cpu(), t.new_zeros(pad).cpu()]) if pad else t.cpu())
        index[key] = (which, cur[which], t.numel())
        cur[which] += t.numel() + pad

    for _, key, is32 in _TOP_FIELDS:
        add(key, "f32" if is32 else "mm")
    for prefix, layers in (("visual.transformer", arch.v_layers), ("transformer", arch.t_layers)):
        for i in range(layers):
            for _, suffix, is32 in _BLOCK_FIELDS:
                add(f"{prefix}.resblocks.{i}.{suffix}", "f32" if is32 else "mm")
    return torch.cat(mm), torch.cat(f32), index
