"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the reference hot path (the oracle).

Nothing in the product path (rpo_b200/, trainers/) may import this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs do, and only as the
checker or the timed CPU baseline.

What it restates (all citations into /root/reference):
  * trainers/rpo.py:132-138  make_prompts  (token embedding + positional embedding, len_prompts)
  * trainers/rpo.py:140-159  define_mask   (dense additive masks, reference-shaped)
  * trainers/rpo.py:161-232  CustomCLIP.forward (text splice, both towers, K-pair logits, CE)
  * clip/model.py:153-159    LayerNorm (fp32 compute, cast back)
  * clip/model.py:162-164    QuickGELU
  * clip/model.py:181-191    ResidualAttentionBlock (nn.MultiheadAttention with additive float mask)
  * clip/model.py:202-207    Transformer
  * clip/model.py:379-400    convert_weights (which tensors become fp16)

The arithmetic itself lives in PyTorch (unpinned by the reference: requirements.txt:1-3); the
effective oracle version is this image's torch 2.11.0.  The restatement issues the same torch ops
in the same order as the reference so that on the same device it is bit-identical to it
(tests/test_oracle_vs_reference.py checks exactly that where /root/reference is mounted, and
tests/test_oracle_golden.py checks it against tests/golden/*.npz generated from the unmodified
reference by oracle/make_golden.py).  Parity pin: the reference ships no tests or golden vectors of
its own, so the pins are outputs of the reference itself run in the authoring
container.

Parameterised where the reference hard-codes ViT-B/16 constants (trainers/rpo.py:52,142,154,185:
d_v=768, attn_head=8, 1+14*14, 512) so ViT-L/14 and bf16 (BASELINE.json config 3) have an oracle;
at ViT-B/16 the parameterised values equal the hard-coded ones.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

_HALF_KEYS_SUFFIX = (
    "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
    "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias",
)
_HALF_KEYS_EXACT = ("visual.conv1.weight", "text_projection", "visual.proj")

PREC_DTYPE = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16, "fp64": torch.float64}


def convert_state_dict(sd, prec: str, values_of: str = None):
    """clip/model.py:379-400: Linear/Conv/MHA weights+biases and the two projections become 16-bit;
    LayerNorm params, embeddings, class_embedding, logit_scale stay fp32.  bf16 is an
    extension with the same key set (H9).

    prec="fp64" is the TRUTH model of the parity tests (no reference counterpart): the weight VALUES of
    the `values_of` precision (what a fp16 / bf16 / fp32 model holds) carried in double, so that the same
    function is evaluated with (nearly) exact arithmetic."""
    if prec == "fp64":
        src = convert_state_dict(sd, values_of or "fp32")
        return OrderedDict((k, v.double() if v.is_floating_point() else v) for k, v in src.items())
    dt = PREC_DTYPE[prec]
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith(_HALF_KEYS_SUFFIX) or k in _HALF_KEYS_EXACT:
            # build_model always goes through fp16 (clip/model.py:438); PREC=fp32 then calls
            # clip_model.float() (trainers/rpo.py:247-249), i.e. fp16-rounded values held in fp32.
            out[k] = v.to(torch.float16).float() if dt == torch.float32 else v.to(dt)
        else:
            out[k] = v.float() if v.is_floating_point() else v
    return out


def layer_norm(x, w, b):
    # clip/model.py:156-159
    orig = x.dtype
    if orig == torch.float64:  # truth model: no fp32 detour
        return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
    ret = F.layer_norm(x.type(torch.float32), (x.shape[-1],), w, b, 1e-5)
    return ret.type(orig)


def quick_gelu(x):
    # clip/model.py:164
    return x * torch.sigmoid(1.702 * x)


def res_block(x, sd, prefix, n_head, attn_mask):
    # clip/model.py:181-191; nn.MultiheadAttention.forward -> F.multi_head_attention_forward
    D = x.shape[-1]
    h = layer_norm(x, sd[f"{prefix}.ln_1.weight"], sd[f"{prefix}.ln_1.bias"])
    mask = attn_mask.to(dtype=x.dtype, device=x.device)
    a = F.multi_head_attention_forward(
        h, h, h, D, n_head,
        sd[f"{prefix}.attn.in_proj_weight"], sd[f"{prefix}.attn.in_proj_bias"],
        None, None, False, 0.0,
        sd[f"{prefix}.attn.out_proj.weight"], sd[f"{prefix}.attn.out_proj.bias"],
        training=False, need_weights=False, attn_mask=mask)[0]
    x = x + a
    h = layer_norm(x, sd[f"{prefix}.ln_2.weight"], sd[f"{prefix}.ln_2.bias"])
    h = F.linear(h, sd[f"{prefix}.mlp.c_fc.weight"], sd[f"{prefix}.mlp.c_fc.bias"])
    h = quick_gelu(h)
    h = F.linear(h, sd[f"{prefix}.mlp.c_proj.weight"], sd[f"{prefix}.mlp.c_proj.bias"])
    return x + h


def transformer(x, sd, prefix, layers, n_head, attn_mask, taps=None):
    # clip/model.py:202-207
    for i in range(layers):
        x = res_block(x, sd, f"{prefix}.resblocks.{i}", n_head, attn_mask)
        if taps is not None:
            taps.append(x)
    return x


def n_layers(sd, prefix):
    return len({k.split(".resblocks.")[1].split(".")[0] for k in sd if k.startswith(prefix + ".resblocks.")})


def make_text_x(sd, tokens, dtype):
    """trainers/rpo.py:135-137.  Returns (text_x [C,T,Dt] in model dtype, len_prompts [C])."""
    with torch.no_grad():
        emb = sd["token_embedding.weight"][tokens]
        text_x = emb.type(dtype) + sd["positional_embedding"].type(dtype)
        len_prompts = tokens.argmax(dim=-1) + 1
    return text_x, len_prompts


def define_masks(len_prompts, K, n_ctx_v, heads_t, dtype, len_max=77):
    """trainers/rpo.py:140-159.  text_mask [C*heads_t, T, T] fp32, visual_mask [S+K, S+K] dtype."""
    masks = []
    for idx in len_prompts.tolist():
        m = torch.empty(len_max, len_max)
        m.fill_(float("-inf"))
        m.triu_(1)
        m[:, idx:].fill_(float("-inf"))
        masks.append(m.repeat(heads_t, 1, 1))
    text_mask = torch.cat(masks) if masks else torch.empty(0, len_max, len_max)
    att = n_ctx_v + K
    visual_mask = torch.zeros((att, att), dtype=dtype)
    visual_mask[:, -K:] = float("-inf")
    return text_mask, visual_mask


class OracleModel:
    """Holds what CustomCLIP.__init__ precomputes (trainers/rpo.py:99-130).  `sd` must already be in
    the model precision (convert_state_dict)."""

    def __init__(self, sd, tokens, K, prec="fp32", device="cpu"):
        self.prec = prec
        self.dtype = PREC_DTYPE[prec]
        self.device = torch.device(device)
        self.sd = {k: v.to(self.device) for k, v in sd.items()}
        self.K = K
        self.tokens = tokens
        Dv = sd["visual.class_embedding"].shape[0]
        Dt = sd["ln_final.weight"].shape[0]
        self.heads_v = Dv // 64
        self.heads_t = Dt // 64  # == 8 at ViT-B/16 (trainers/rpo.py:142)
        self.S = sd["visual.positional_embedding"].shape[0]  # == 1 + 14*14 at ViT-B/16 (:154)
        self.layers_v = n_layers(sd, "visual.transformer")
        self.layers_t = n_layers(sd, "transformer")
        self.text_x, self.len_prompts = make_text_x(sd, tokens, self.dtype)
        self.text_mask, self.visual_mask = define_masks(self.len_prompts, K, self.S, self.heads_t, self.dtype,
                                                        len_max=tokens.shape[1])
        # the reference keeps these on the host and re-uploads them; the oracle may
        # keep them resident -- same values either way.
        self.text_mask = self.text_mask.to(self.device)
        self.visual_mask = self.visual_mask.to(self.device)

    def forward(self, image, text_prompt, img_prompt, label=None, training=True, taps=None):
        """trainers/rpo.py:161-232.  Returns the scalar CE loss if `training` else logits [B,C] fp32.
        `taps`, if a dict, receives intermediate tensors for layer-wise parity checks."""
        sd, K, dtype, device = self.sd, self.K, self.dtype, self.device
        C = self.text_x.shape[0]
        ar = torch.arange(C)
        # ---- text ---- (:173-192)
        text_x = self.text_x.to(device).clone()
        for i in range(K):
            text_x[ar, self.len_prompts + i, :] = text_prompt[i, :].repeat(C, 1)
        text_x = text_x.permute(1, 0, 2)
        ttaps = [] if taps is not None else None
        text_x = transformer(text_x, sd, "transformer", self.layers_t, self.heads_t, self.text_mask, ttaps)
        text_x = text_x.permute(1, 0, 2)
        text_x = layer_norm(text_x, sd["ln_final.weight"], sd["ln_final.bias"]).type(dtype)
        text_f = torch.empty(C, 0, text_x.shape[-1], device=device, dtype=dtype)
        for i in range(K):
            idx = self.len_prompts + i
            x = text_x[ar, idx]
            text_f = torch.cat([text_f, x[:, None, :]], dim=1)
        text_f = text_f @ sd["text_projection"]
        # ---- image ---- (:195-211)
        B = image.shape[0]
        emb = F.conv2d(image.type(dtype), sd["visual.conv1.weight"], None,
                       stride=sd["visual.conv1.weight"].shape[-1])
        emb = emb.reshape(B, emb.shape[1], -1).permute(0, 2, 1)
        emb = torch.cat([sd["visual.class_embedding"].repeat(B, 1, 1).type(dtype), emb], dim=1)
        img_x = emb + sd["visual.positional_embedding"].type(dtype)
        img_x = torch.cat([img_x, img_prompt.repeat(B, 1, 1)], dim=1)
        img_x = layer_norm(img_x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
        if taps is not None:
            taps["img_x0"] = img_x
        img_x = img_x.permute(1, 0, 2)
        vtaps = [] if taps is not None else None
        img_x = transformer(img_x, sd, "visual.transformer", self.layers_v, self.heads_v, self.visual_mask, vtaps)
        img_x = img_x.permute(1, 0, 2)
        img_f = layer_norm(img_x[:, -K:, :], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]) @ sd["visual.proj"]
        # ---- logits ---- (:215-227)
        text_f = text_f / text_f.norm(dim=-1, keepdim=True)
        img_f = img_f / img_f.norm(dim=-1, keepdim=True)
        logits = torch.zeros(B, C, device=device, dtype=torch.float64 if dtype == torch.float64 else torch.float32)
        for i in range(K):
            logit = sd["logit_scale"].exp() * img_f[:, i, :] @ text_f[:, i, :].t()
            logits += logit
        logits /= K
        if taps is not None:
            taps["text_layers"] = [t.permute(1, 0, 2) for t in ttaps]
            taps["img_layers"] = [t.permute(1, 0, 2) for t in vtaps]
            taps["text_f"] = text_f
            taps["img_f"] = img_f
            taps["logits"] = logits
        if training:
            return F.cross_entropy(logits, label.to(device))
        return logits

    def step(self, image, text_prompt, img_prompt, label):
        """forward + CE + backward to the two prompt gradients (trainers/rpo.py:306-308)."""
        tp = text_prompt.detach().to(self.device, self.dtype).requires_grad_(True)
        ip = img_prompt.detach().to(self.device, self.dtype).requires_grad_(True)
        loss = self.forward(image.to(self.device), tp, ip, label, training=True)
        loss.backward()
        return loss.detach(), tp.grad.detach(), ip.grad.detach()

    def logits(self, image, text_prompt, img_prompt, taps=None):
        with torch.no_grad():
            return self.forward(image.to(self.device), text_prompt.to(self.device, self.dtype),
                                img_prompt.to(self.device, self.dtype), None, training=False, taps=taps)
