"""Synthetic CLIP weights, tokens and inputs (there is no network for real checkpoints).

`make_state_dict(arch, seed)` produces a state dict with exactly the key set and shapes that
`clip.model.build_model` infers its hyper-parameters from (reference clip/model.py:403-440), so the
same dict can be loaded into the unmodified reference (oracle/ref_harness.py), fed to the oracle
restatement (oracle/rpo_oracle.py) and bound to the CUDA path (rpo_b200.model).  Initialisation
scales follow clip/model.py:303-330; LayerNorm affine parameters and all biases are perturbed by
0.1*N(0,1) so the LN/bias code paths are exercised.  All randomness comes from one
CPU `torch.Generator`, which is deterministic for a given torch build.
"""
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class ClipArch:
    name: str
    embed_dim: int
    image_resolution: int
    vision_layers: int
    vision_width: int
    vision_patch_size: int
    context_length: int
    vocab_size: int
    transformer_width: int
    transformer_heads: int
    transformer_layers: int

    @property
    def vision_heads(self) -> int:
        return self.vision_width // 64

    @property
    def grid(self) -> int:
        return self.image_resolution // self.vision_patch_size

    @property
    def n_patch(self) -> int:
        return self.grid * self.grid


ARCHS = {
    # reference clip/clip.py:29-36 "ViT-B/16"
    "ViT-B/16": ClipArch("ViT-B/16", 512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    # not in the reference's _MODELS; BASELINE.json config 3
    "ViT-L/14": ClipArch("ViT-L/14", 768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    # small shapes for fast parity tests: same structure, 2 layers
    "tiny": ClipArch("tiny", 128, 64, 2, 128, 16, 77, 49408, 128, 2, 2),
    "small": ClipArch("small", 256, 96, 3, 256, 16, 77, 49408, 192, 3, 3),
}


def make_state_dict(arch: ClipArch, seed: int = 0, perturb: float = 0.1):
    g = torch.Generator().manual_seed(seed)

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {}
    Dv, Dt, E = arch.vision_width, arch.transformer_width, arch.embed_dim
    p = arch.vision_patch_size

    def block(prefix, width, layers):
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        for i in range(layers):
            b = f"{prefix}.resblocks.{i}"
            sd[f"{b}.attn.in_proj_weight"] = randn(3 * width, width, std=attn_std)
            sd[f"{b}.attn.in_proj_bias"] = randn(3 * width, std=perturb)
            sd[f"{b}.attn.out_proj.weight"] = randn(width, width, std=proj_std)
            sd[f"{b}.attn.out_proj.bias"] = randn(width, std=perturb)
            sd[f"{b}.ln_1.weight"] = 1.0 + randn(width, std=perturb)
            sd[f"{b}.ln_1.bias"] = randn(width, std=perturb)
            sd[f"{b}.mlp.c_fc.weight"] = randn(4 * width, width, std=fc_std)
            sd[f"{b}.mlp.c_fc.bias"] = randn(4 * width, std=perturb)
            sd[f"{b}.mlp.c_proj.weight"] = randn(width, 4 * width, std=proj_std)
            sd[f"{b}.mlp.c_proj.bias"] = randn(width, std=perturb)
            sd[f"{b}.ln_2.weight"] = 1.0 + randn(width, std=perturb)
            sd[f"{b}.ln_2.bias"] = randn(width, std=perturb)

    scale = Dv ** -0.5
    sd["visual.class_embedding"] = randn(Dv, std=scale)
    sd["visual.positional_embedding"] = randn(arch.n_patch + 1, Dv, std=scale)
    sd["visual.proj"] = randn(Dv, E, std=scale)
    sd["visual.conv1.weight"] = randn(Dv, 3, p, p, std=(3 * p * p) ** -0.5)
    sd["visual.ln_pre.weight"] = 1.0 + randn(Dv, std=perturb)
    sd["visual.ln_pre.bias"] = randn(Dv, std=perturb)
    sd["visual.ln_post.weight"] = 1.0 + randn(Dv, std=perturb)
    sd["visual.ln_post.bias"] = randn(Dv, std=perturb)
    block("visual.transformer", Dv, arch.vision_layers)

    sd["token_embedding.weight"] = randn(arch.vocab_size, Dt, std=0.02)
    sd["positional_embedding"] = randn(arch.context_length, Dt, std=0.01)
    sd["ln_final.weight"] = 1.0 + randn(Dt, std=perturb)
    sd["ln_final.bias"] = randn(Dt, std=perturb)
    sd["text_projection"] = randn(Dt, E, std=Dt ** -0.5)
    sd["logit_scale"] = torch.tensor(2.6592600369327783)  # log(1/0.07), clip/model.py:299
    block("transformer", Dt, arch.transformer_layers)
    return sd


def make_images(batch: int, resolution: int = 224, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, resolution, resolution, generator=g)


def make_labels(batch: int, n_cls: int):
    return torch.arange(batch, dtype=torch.int64) % n_cls


def make_prompt_init(state_dict, K: int, seed: int = 7):
    """Same recipe as PromptLearner.initialization_token (reference trainers/rpo.py:60-88) but with a
    private generator, so tests do not depend on the global RNG: EOT-token embedding / class
    embedding repeated K times plus 0.1 * unit-norm Gaussian noise.  Returned in fp32."""
    g = torch.Generator().manual_seed(seed)
    Dt = state_dict["ln_final.weight"].shape[0]
    Dv = state_dict["visual.class_embedding"].shape[0]
    tn = torch.randn(K, Dt, generator=g)
    tn = tn / tn.norm(dim=-1, keepdim=True)
    text_prompt = state_dict["token_embedding.weight"][49407].repeat(K, 1) + 0.1 * tn
    vn = torch.randn(K, Dv, generator=g)
    vn = vn / vn.norm(dim=-1, keepdim=True)
    img_prompt = state_dict["visual.class_embedding"].repeat(K, 1) + 0.1 * vn
    return text_prompt, img_prompt


def synthetic_classnames(n_cls: int):
    return [f"class {i}" for i in range(n_cls)]
