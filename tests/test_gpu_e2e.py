"""GPU parity of the whole hot path through the reference-shaped surface (rpo_b200.model.CustomCLIP ->
C ABI -> sm_100a kernels) against (a) the golden vectors generated from the UNMODIFIED reference
(tests/golden/*.npz, oracle/make_golden.py) and (b) the oracle restatement run on the same seeded
inputs.  Tolerances (BASELINE.json north_star): 1e-5 for fp32, 1e-3 for fp16 -- applied to the loss
absolutely and to logits / activations relative to the tensor's max magnitude.  bf16 (an extension,
SURVEY H9) has an 8-bit mantissa: 1e-2.  Gradients: see GRAD_TRUTH_TOL / FLOOR / DIRECT below -- they are
measured against a float64 evaluation of the same function, and every error is printed (pytest -s).
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.rpo_oracle import OracleModel, convert_state_dict
from rpo_b200 import _lib, synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from tests.common import GOLDEN_CASES, check_grads, class_tokens, load_golden, rel_err, state_dict, truth_grads

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-5, "fp16": 1e-3, "bf16": 1e-2}
def make_cfg(K, res):
    return SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC="fp16")),
                           INPUT=SimpleNamespace(SIZE=(res, res)))


def build_model(arch_name, prec, K, tokens, backend=_lib.GEMM_AUTO, seed=0):
    arch = synth.ARCHS[arch_name]
    sd = state_dict(arch_name, seed)
    clip = SyntheticCLIP(sd, prec)
    names = [f"c{i}" for i in range(tokens.shape[0])]
    model = CustomCLIP(make_cfg(K, arch.image_resolution), names, "a photo of a _.", clip, tokens=tokens,
                       gemm_backend=backend)
    return model.to("cuda:0"), arch, sd


def set_prompts(model, tp, ip):
    with torch.no_grad():
        model.prompt_learner.text_prompt.copy_(tp.to(model.dtype))
        model.prompt_learner.img_prompt.copy_(ip.to(model.dtype))


def step(model, image, label):
    model.prompt_learner.train()
    for p in model.prompt_learner.parameters():
        p.grad = None
    loss = model(image, label)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), model.prompt_learner.text_prompt.grad.float().cpu(), \
        model.prompt_learner.img_prompt.grad.float().cpu()


def eval_logits(model, image):
    model.prompt_learner.eval()
    with torch.no_grad():
        out = model(image)
    model.prompt_learner.train()
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_matches_reference_golden(name):
    """The CUDA path reproduces what the unmodified reference produced for the same seeded inputs."""
    g = load_golden(name)
    prec = name.rsplit("_", 1)[1]
    K, B = int(g["K"]), int(g["B"])
    tokens = torch.from_numpy(g["tokens"].astype(np.int64))
    model, arch, sd = build_model("ViT-B/16", prec, K, tokens)
    set_prompts(model, torch.from_numpy(g["text_prompt"]), torch.from_numpy(g["img_prompt"]))
    image = synth.make_images(B, arch.image_resolution).cuda()
    label = synth.make_labels(B, tokens.shape[0]).cuda()
    loss, gt, gi = step(model, image, label)
    logits = eval_logits(model, image)
    tol = TOL[prec]
    scale = float(np.exp(2.6592600369327783))  # logits are exp(logit_scale) * cosine
    print(f"{name}: dloss={abs(loss.item() - float(g['loss'])):.3e} "
          f"dlogits={(logits - torch.from_numpy(g['logits'])).abs().max().item():.3e} "
          f"gt={rel_err(gt, torch.from_numpy(g['grad_text_prompt'])):.3e} "
          f"gi={rel_err(gi, torch.from_numpy(g['grad_img_prompt'])):.3e}")
    assert abs(loss.item() - float(g["loss"])) <= tol * max(1.0, abs(float(g["loss"])))
    assert (logits - torch.from_numpy(g["logits"])).abs().max().item() <= tol * scale
    tgt, tgi = truth_grads(sd, prec, tokens, K, image, model.prompt_learner.text_prompt.detach(),
                           model.prompt_learner.img_prompt.detach(), label)
    check_grads(prec, gt, torch.from_numpy(g["grad_text_prompt"]), tgt, "text")
    check_grads(prec, gi, torch.from_numpy(g["grad_img_prompt"]), tgi, "image")
    # residual-stream rows after every block (forward hooks in the reference)
    eng = model._engine
    S = arch.n_patch + 1
    lp = model.len_prompts
    off0 = 0  # class 0 owns the first len_prompts[0] context rows
    Mc_t = int(lp.sum())
    for layer in range(arch.vision_layers):
        x = eng.debug_fetch(0, layer).float().cpu()  # [B*S ctx rows | B*K prompt rows]
        rows = []
        for r in g["rows_v"].tolist():
            rows.append(x[r] if r < S else x[B * S + (r - S)])  # image 0
        got = torch.stack(rows)
        assert rel_err(got, torch.from_numpy(g["taps_v"][layer])) <= tol * 5, f"vision block {layer}"
    for layer in range(arch.transformer_layers):
        x = eng.debug_fetch(1, layer).float().cpu()
        rows = []
        for r in g["rows_t"].tolist():
            rows.append(x[off0 + r] if r < int(lp[0]) else x[Mc_t + (r - int(lp[0]))])  # class 0
        got = torch.stack(rows)
        assert rel_err(got, torch.from_numpy(g["taps_t"][layer])) <= tol * 5, f"text block {layer}"


E2E_CASES = [
    # arch, prec, K, class ids, B
    ("tiny", "fp32", 5, [3, 77, 512], 2),
    ("tiny", "fp16", 5, [3, 77, 512], 2),
    ("tiny", "bf16", 5, [3, 77, 512], 2),
    ("small", "fp32", 8, [0, 10, 100, 999], 5),
    ("small", "fp16", 8, [0, 10, 100, 999], 5),
    ("ViT-B/16", "fp16", 24, list(range(0, 1000, 53)), 4),
    # BASELINE config 3 geometry (ViT-L/14, 257 context rows + 24 prompts, bf16) at a small batch: 588 -> 640 padded
    # patch GEMM, two key blocks in the tcgen05 attention, 24-layer towers (full size: test_full_size_configs)
    ("ViT-L/14", "bf16", 24, [1, 20, 300], 2),
]


@pytest.mark.parametrize("arch_name,prec,K,class_ids,B", E2E_CASES)
@pytest.mark.parametrize("backend", [_lib.GEMM_AUTO, _lib.GEMM_SIMT], ids=["auto", "simt"])
def test_matches_oracle(arch_name, prec, K, class_ids, B, backend):
    """Same seeded inputs through the CUDA path and through the oracle (torch, fp32/fp16 on the GPU so
    the 16-bit arithmetic is the reference's own on this device)."""
    if arch_name == "ViT-B/16" and backend == _lib.GEMM_SIMT:
        pytest.skip("covered by the golden test")
    tokens = class_tokens(class_ids)
    model, arch, sd = build_model(arch_name, prec, K, tokens, backend)
    tp, ip = synth.make_prompt_init(sd, K)
    set_prompts(model, tp, ip)
    image = synth.make_images(B, arch.image_resolution)
    label = synth.make_labels(B, len(class_ids))
    loss, gt, gi = step(model, image.cuda(), label.cuda())
    logits = eval_logits(model, image.cuda())
    om = OracleModel(convert_state_dict(sd, prec), tokens, K, prec, device="cuda:0")
    tpd = model.prompt_learner.text_prompt.detach()
    ipd = model.prompt_learner.img_prompt.detach()
    oloss, ogt, ogi = om.step(image, tpd, ipd, label)
    ologits = om.logits(image, tpd, ipd)
    tol = TOL[prec]
    scale = float(np.exp(2.6592600369327783))
    print(f"{arch_name}/{prec}: dloss={abs(loss.item() - oloss.item()):.3e} "
          f"dlogits={(logits - ologits.cpu()).abs().max().item():.3e} gt={rel_err(gt, ogt):.3e} "
          f"gi={rel_err(gi, ogi):.3e}")
    assert abs(loss.item() - oloss.item()) <= tol * max(1.0, abs(oloss.item()))
    assert (logits - ologits.cpu()).abs().max().item() <= tol * scale
    del om
    tgt, tgi = truth_grads(sd, prec, tokens, K, image, tpd, ipd, label)
    check_grads(prec, gt, ogt, tgt, "text")
    check_grads(prec, gi, ogi, tgi, "image")


def test_config2_full_size_properties():
    """BASELINE config 2 (ViT-B/16, K=24, 100 classes, batch 32, fp16) at full size: compared with the
    oracle on the GPU, plus size-independent properties -- non-prompt rows do not depend on the
    prompts (read-only), and logits are invariant to the order of images in the batch."""
    K, Cn, B = 24, 100, 32
    tokens = class_tokens(range(Cn))
    model, arch, sd = build_model("ViT-B/16", "fp16", K, tokens)
    tp, ip = synth.make_prompt_init(sd, K)
    set_prompts(model, tp, ip)
    image = synth.make_images(B, arch.image_resolution).cuda()
    label = synth.make_labels(B, Cn).cuda()
    loss, gt, gi = step(model, image, label)
    logits = eval_logits(model, image)
    eng = model._engine
    S = arch.n_patch + 1
    ctx_before = eng.debug_fetch(0, arch.vision_layers - 1)[:B * S].clone()
    om = OracleModel(convert_state_dict(sd, "fp16"), tokens, K, "fp16", device="cuda:0")
    tpd, ipd = model.prompt_learner.text_prompt.detach(), model.prompt_learner.img_prompt.detach()
    oloss, ogt, ogi = om.step(image, tpd, ipd, label)
    ologits = om.logits(image, tpd, ipd).cpu()
    scale = float(np.exp(2.6592600369327783))
    print(f"cfg2: loss {loss.item():.5f} vs {oloss.item():.5f}; dlogits={(logits - ologits).abs().max().item():.3e} "
          f"gt={rel_err(gt, ogt):.3e} gi={rel_err(gi, ogi):.3e}")
    assert abs(loss.item() - oloss.item()) <= 1e-3 * max(1.0, abs(oloss.item()))
    assert (logits - ologits).abs().max().item() <= 1e-3 * scale
    del om
    tgt, tgi = truth_grads(sd, "fp16", tokens, K, image, tpd, ipd, label)
    check_grads("fp16", gt, ogt, tgt, "text")
    check_grads("fp16", gi, ogi, tgi, "image")
    # permutation of the batch permutes the logits rows, bit-exactly
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    logits_p = eval_logits(model, image[perm.cuda()])
    assert torch.equal(logits_p, logits[perm])
    # read-only: perturbing the prompts leaves every context row of the last block bit-identical
    set_prompts(model, tp + 0.5, ip - 0.25)
    eval_logits(model, image)
    ctx_after = eng.debug_fetch(0, arch.vision_layers - 1)[:B * S]
    assert torch.equal(ctx_before, ctx_after)


def test_errors_are_loud():
    tokens = class_tokens([1, 2])
    model, arch, _ = build_model("tiny", "fp16", 4, tokens)
    with pytest.raises(_lib.RpoError):
        model(torch.zeros(1, 3, arch.image_resolution, arch.image_resolution))  # CPU tensor: no fallback
    with pytest.raises(_lib.RpoError):
        model(torch.zeros(1, 3, 32, 32, device="cuda:0"), torch.zeros(1, dtype=torch.int64, device="cuda:0"))
    with pytest.raises(IndexError):  # prompt + K does not fit in 77 tokens (trainers/rpo.py:177)
        long_tokens = tokens.clone()
        long_tokens[0, 1:76] = 320  # push the EOT token (the argmax) to the last position
        long_tokens[0, 76] = 49407
        build_model("tiny", "fp16", 4, long_tokens)


@pytest.mark.parametrize("arch_name,prec", [("tiny", "fp16"), ("ViT-B/16", "fp16"), ("tiny", "fp32"), ("ViT-B/16", "bf16")])
def test_uint8_images_match_the_float_pipeline(arch_name, prec):
    """raw uint8 pixels, ToTensor + Normalize (clip/clip.py:75-78) applied inside the patch extraction,
    against the same preprocessing done on the host in IEEE f32 (numpy: true divisions) and uploaded as float32 --
    the reference's data path (DataLoader tensor -> `.to(device)` -> `image.type(self.dtype)`).  Bit-identical."""
    tokens = class_tokens([3, 14, 159])
    model, arch, _ = build_model(arch_name, prec, 4, tokens)
    res = arch.image_resolution
    rng = np.random.default_rng(5)
    img_u8 = rng.integers(0, 256, size=(3, 3, res, res), dtype=np.uint8)
    img_u8[0, :, 0, :8] = [[0, 1, 2, 127, 128, 253, 254, 255]] * 3  # range ends
    mean = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32).reshape(1, 3, 1, 1)
    std = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32).reshape(1, 3, 1, 1)
    img_f32 = (img_u8.astype(np.float32) / np.float32(255.0) - mean) / std
    assert img_f32.dtype == np.float32
    a = eval_logits(model, torch.from_numpy(img_u8).to("cuda:0"))
    b = eval_logits(model, torch.from_numpy(img_f32).to("cuda:0"))
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
    # other constants go through rpo_set_image_norm
    model.engine(3).set_image_norm([0.5, 0.5, 0.5], [0.25, 0.5, 1.0])
    m2 = np.array([0.5, 0.5, 0.5], dtype=np.float32).reshape(1, 3, 1, 1)
    s2 = np.array([0.25, 0.5, 1.0], dtype=np.float32).reshape(1, 3, 1, 1)
    c = eval_logits(model, torch.from_numpy(img_u8).to("cuda:0"))
    d = eval_logits(model, torch.from_numpy((img_u8.astype(np.float32) / np.float32(255.0) - m2) / s2).to("cuda:0"))
    assert torch.equal(c, d) and not torch.equal(a, c)
    # training step from uint8 pixels
    label = torch.tensor([0, 1, 2], device="cuda:0")
    l8, gt8, gi8 = step(model, torch.from_numpy(img_u8).to("cuda:0"), label)
    lf, gtf, gif = step(model, torch.from_numpy((img_u8.astype(np.float32) / np.float32(255.0) - m2) / s2).to("cuda:0"), label)
    assert torch.equal(l8, lf) and torch.equal(gt8, gtf) and torch.equal(gi8, gif)


def test_eval_reuses_text_features():
    """at test time the text features depend on the prompts only; they are computed once and reused
    until the prompts change (the reference reruns the text tower per batch, trainers/rpo.py:173-192)."""
    tokens = class_tokens([1, 20, 300, 4])
    model, arch, _ = build_model("tiny", "fp16", 4, tokens)
    res = arch.image_resolution
    g = torch.Generator().manual_seed(3)
    img1 = torch.randn(2, 3, res, res, generator=g).to("cuda:0")
    img2 = torch.randn(3, 3, res, res, generator=g).to("cuda:0")
    eng = model.engine(3)
    first = eval_logits(model, img1)
    n_full = eng.launch_count()
    cached = eval_logits(model, img2)          # second batch: text tower skipped
    n_cached = eng.launch_count()
    assert n_cached < n_full
    model._text_key = None
    fresh = eval_logits(model, img2)
    assert eng.launch_count() == n_full
    assert torch.equal(cached, fresh)
    # the cache follows the prompts: in-place edit, training-mode forward, fused optimiser step
    with torch.no_grad():
        model.prompt_learner.text_prompt.add_(0.05)
    moved = eval_logits(model, img2)
    assert eng.launch_count() == n_full and not torch.equal(moved, fresh)
    again = eval_logits(model, img2)
    assert eng.launch_count() == n_cached and torch.equal(again, moved)
    step(model, img1, torch.tensor([0, 1], device="cuda:0"))
    eval_logits(model, img2)
    n_recomputed = eng.launch_count()          # (the count now also holds the backward launches of the step)
    eval_logits(model, img2)
    assert n_recomputed - eng.launch_count() == n_full - n_cached
    # C ABI contract: cached features are for inference only
    with pytest.raises(_lib.RpoError):
        eng.forward(img1, None, model.prompt_learner.img_prompt.data, torch.tensor([0, 1], device="cuda:0"))
    assert first.shape == (2, 4)


def _long_tokens(n_words, width=77):
    """A prompt of `n_words` tokens between SOT (49406) and EOT (49407): len_prompts = n_words + 2."""
    t = torch.zeros(1, width, dtype=torch.int64)
    t[0, 0] = 49406
    t[0, 1:1 + n_words] = torch.arange(n_words) % 1000 + 320
    t[0, 1 + n_words] = 49407
    return t


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("case", ["k1", "k1_b1", "one_class_one_image", "max_length", "ragged_with_max"])
def test_edge_shapes_match_oracle(prec, case):
    """Edges of the reference's index arithmetic (trainers/rpo.py:137,149,177): a single prompt pair, a single
    class / image, a class prompt whose K prompt slots end exactly at position 76 (len_prompts + K == 77), and a
    class list mixing the shortest and the longest admissible prompts."""
    K, B = {"k1": (1, 3), "k1_b1": (1, 1), "one_class_one_image": (3, 1), "max_length": (4, 2),
            "ragged_with_max": (4, 3)}[case]
    if case in ("k1", "k1_b1"):  # k1_b1: one pair, one image -- the largest dlogits the fp16 backward can see
        tokens = class_tokens([5, 600])
    elif case == "one_class_one_image":
        tokens = class_tokens([42])
    elif case == "max_length":
        tokens = _long_tokens(77 - K - 2)  # len_prompts = 77 - K
    else:
        tokens = torch.cat([class_tokens([7]), _long_tokens(77 - K - 2), _long_tokens(1), class_tokens([999])])
    model, arch, sd = build_model("tiny", prec, K, tokens)
    assert int(model.len_prompts.max()) + K <= 77
    tp, ip = synth.make_prompt_init(sd, K)
    set_prompts(model, tp, ip)
    Cn = tokens.shape[0]
    image = synth.make_images(B, arch.image_resolution)
    label = synth.make_labels(B, Cn)
    loss, gt, gi = step(model, image.cuda(), label.cuda())
    logits = eval_logits(model, image.cuda())
    om = OracleModel(convert_state_dict(sd, prec), tokens, K, prec, device="cuda:0")
    tpd, ipd = model.prompt_learner.text_prompt.detach(), model.prompt_learner.img_prompt.detach()
    oloss, ogt, ogi = om.step(image, tpd, ipd, label)
    ologits = om.logits(image, tpd, ipd).cpu()
    tol = TOL[prec]
    assert logits.shape == (B, Cn)
    assert abs(loss.item() - oloss.item()) <= tol * max(1.0, abs(oloss.item()))
    assert (logits - ologits).abs().max().item() <= tol * float(np.exp(2.6592600369327783))
    if Cn == 1:
        # one class: softmax over a single logit, the loss is exactly 0 and so are both gradients
        assert loss.item() == 0.0 and float(gt.abs().max()) == 0.0 and float(gi.abs().max()) == 0.0
        assert float(ogt.abs().max()) == 0.0
        return
    del om
    tgt, tgi = truth_grads(sd, prec, tokens, K, image, tpd, ipd, label)
    check_grads(prec, gt, ogt, tgt, "text")
    check_grads(prec, gi, ogi, tgi, "image")


# BASELINE.json configs 3, 4, 5 at their FULL sizes; config 2 is test_config2_full_size_properties.
# The 1000-class shapes use the fp32 oracle as the truth (see truth_grads).
FULL_SIZE = [
    # id, arch, prec, K, classes, batch
    ("cfg3_vitl14_bf16_b16", "ViT-L/14", "bf16", 24, 100, 16),
    ("cfg4_c1000_b32", "ViT-B/16", "fp16", 24, 1000, 32),
    ("cfg5_k4_c1000_b64", "ViT-B/16", "fp16", 4, 1000, 64),
    ("cfg5_k8_c1000_b64", "ViT-B/16", "fp16", 8, 1000, 64),
    ("cfg5_k16_c1000_b64", "ViT-B/16", "fp16", 16, 1000, 64),
    ("cfg5_k48_c1000_b64", "ViT-B/16", "fp16", 48, 1000, 64),
]


@pytest.mark.parametrize("name,arch_name,prec,K,Cn,B", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_full_size_configs(name, arch_name, prec, K, Cn, B):
    """End-to-end parity (loss, eval logits, both prompt gradients) at the sizes BASELINE.json names."""
    import gc
    tokens = class_tokens(range(Cn))
    model, arch, sd = build_model(arch_name, prec, K, tokens)
    tp, ip = synth.make_prompt_init(sd, K)
    set_prompts(model, tp, ip)
    image = synth.make_images(B, arch.image_resolution).cuda()
    label = synth.make_labels(B, Cn).cuda()
    loss, gt, gi = step(model, image, label)
    logits = eval_logits(model, image)
    tpd, ipd = model.prompt_learner.text_prompt.detach().clone(), model.prompt_learner.img_prompt.detach().clone()
    model._engine = None
    del model
    gc.collect()
    torch.cuda.empty_cache()
    om = OracleModel(convert_state_dict(sd, prec), tokens, K, prec, device="cuda:0")
    oloss, ogt, ogi = om.step(image, tpd, ipd, label)
    ologits = om.logits(image, tpd, ipd).cpu()
    oloss, ogt, ogi = oloss.cpu(), ogt.float().cpu(), ogi.float().cpu()
    del om
    gc.collect()
    torch.cuda.empty_cache()
    tol = TOL[prec]
    scale = float(np.exp(2.6592600369327783))
    print(f"{name}: loss {loss.item():.5f} vs {oloss.item():.5f} (d={abs(loss.item() - oloss.item()):.3e}); "
          f"dlogits={(logits - ologits).abs().max().item():.3e} of {scale:.1f}")
    assert torch.isfinite(gt).all() and torch.isfinite(gi).all()
    assert abs(loss.item() - oloss.item()) <= tol * max(1.0, abs(oloss.item()))
    assert (logits - ologits).abs().max().item() <= tol * scale
    tgt, tgi = truth_grads(sd, prec, tokens, K, image, tpd, ipd, label, big=Cn >= 1000)
    check_grads(prec, gt, ogt, tgt, "text")
    check_grads(prec, gi, ogi, tgi, "image")
