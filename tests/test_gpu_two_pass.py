"""GPU parity of the two-pass image forward (include/rpo_b200.h rpo_forward_image_context / _prompts): the context
rows of the vision tower do not depend on the prompts (visual_mask, trainers/rpo.py:155-156), so they may be
computed apart from (and ahead of) the prompt rows without changing any result."""
from types import SimpleNamespace

import pytest
import torch

from oracle.rpo_oracle import OracleModel, convert_state_dict
from rpo_b200 import _lib, synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from tests.common import check_grads, class_tokens, rel_err, state_dict, truth_grads

pytestmark = pytest.mark.gpu


def make_model(arch_name, prec, K, tokens, slots=1):
    arch = synth.ARCHS[arch_name]
    sd = state_dict(arch_name, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    model = CustomCLIP(cfg, [f"c{i}" for i in range(tokens.shape[0])], "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=tokens).to("cuda:0")
    model.pipeline_images(slots)
    tp, ip = synth.make_prompt_init(sd, K)
    with torch.no_grad():
        model.prompt_learner.text_prompt.copy_(tp.to(model.dtype))
        model.prompt_learner.img_prompt.copy_(ip.to(model.dtype))
    return model, arch, sd


@pytest.mark.parametrize("arch_name,prec,K,class_ids,B", [
    ("tiny", "fp32", 5, [3, 77, 512], 3),
    ("small", "fp16", 8, [0, 10, 100, 999], 5),
    ("ViT-B/16", "fp16", 24, list(range(0, 1000, 91)), 4),
    ("ViT-L/14", "bf16", 24, [1, 20, 300], 2),
])
def test_two_pass_image_forward_matches_single_pass_and_oracle(arch_name, prec, K, class_ids, B):
    tokens = class_tokens(class_ids)
    Cn = len(class_ids)
    model, arch, sd = make_model(arch_name, prec, K, tokens, slots=2)
    eng = model.engine(B)
    assert eng.image_slots == 2
    image = synth.make_images(B, arch.image_resolution).cuda()
    other = synth.make_images(B, arch.image_resolution, seed=99).cuda()
    label = synth.make_labels(B, Cn).cuda()
    tp, ip = model.prompt_learner.text_prompt.data, model.prompt_learner.img_prompt.data
    # single pass (slot 0)
    loss1, logits1 = eng.forward(image, tp, ip, label, want_logits=True)
    loss1, logits1 = loss1.clone(), logits1.clone()
    grad1 = eng.backward().clone()
    S = arch.n_patch + 1
    ctx1 = eng.debug_fetch(0, arch.vision_layers - 1)[:B * S].clone()
    # two passes into slot 1, with another batch's context rows landing in slot 0 in between
    idt = _lib.RPO_F32
    eng.image_context(image, idt, 1)
    eng.image_context(other, idt, 0)
    eng.text_forward(tp)
    eng.image_prompts(ip, 1)
    eng.logits_forward(label, eng.logits[:B])
    loss2, logits2 = eng.loss.clone(), eng.logits[:B].clone()
    eng.logits_backward()
    eng.text_backward()
    eng.image_backward()
    grad2 = eng.grad_flat.clone()
    ctx2 = eng.debug_fetch(0, arch.vision_layers - 1)[:B * S].clone()
    torch.cuda.synchronize()
    # context rows: same kernels except that the in-projection GEMM no longer carries the prompt rows
    assert rel_err(ctx2, ctx1) <= (1e-6 if prec == "fp32" else 2e-3)
    tol = {"fp32": 1e-5, "fp16": 1e-3, "bf16": 1e-2}[prec]
    scale = 14.285
    assert abs(loss2.item() - loss1.item()) <= tol * max(1.0, abs(loss1.item()))
    assert (logits2 - logits1).abs().max().item() <= tol * scale
    # the two paths differ only in the attention kernel the prompt rows take and in the GEMM tile their q projection
    # rides in: same roundings, different accumulation order
    # (bf16: each path sits within FLOOR["bf16"] = 3.2e-2 of the truth, so the two may differ by twice that)
    gtol = {"fp32": 1e-5, "fp16": 5e-3, "bf16": 6.4e-2}[prec]
    nt = eng.n_text
    print(f"{arch_name}/{prec}: dloss {abs(loss2.item() - loss1.item()):.2e} dlogits {(logits2 - logits1).abs().max().item():.2e} "
          f"text {rel_err(grad2[:nt], grad1[:nt]):.2e} image {rel_err(grad2[nt:], grad1[nt:]):.2e}")
    assert rel_err(grad2[:nt], grad1[:nt]) <= gtol and rel_err(grad2[nt:], grad1[nt:]) <= gtol
    # and against the oracle
    om = OracleModel(convert_state_dict(sd, prec), tokens, K, prec, device="cuda:0")
    oloss, ogt, ogi = om.step(image, tp, ip, label)
    ologits = om.logits(image, tp, ip)
    assert abs(loss2.item() - oloss.item()) <= tol * max(1.0, abs(oloss.item()))
    assert (logits2 - ologits).abs().max().item() <= tol * scale
    del om
    tgt, tgi = truth_grads(sd, prec, tokens, K, image, tp, ip, label)
    check_grads(prec, grad2[:nt].view(K, -1), ogt, tgt, "text")
    check_grads(prec, grad2[nt:].view(K, -1), ogi, tgi, "image")
    # slot checks of the C ABI
    one, _, _ = make_model("tiny", "fp16", 2, class_tokens([1, 2]))
    e1 = one.engine(1)
    with pytest.raises(_lib.RpoError):
        e1.image_context(torch.zeros(1, 3, 32, 32, device="cuda:0"), idt, 1)
