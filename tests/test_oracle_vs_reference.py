"""CPU, authoring container only: the oracle restatement is bit-identical to the unmodified
reference (imported from /root/reference with ftfy/dassl stubs).  Skipped where the reference tree
is not mounted (the GPU box) -- there tests/test_oracle_golden.py carries the pin."""
import pytest
import torch

from oracle import ref_harness as rh
from oracle.rpo_oracle import OracleModel, convert_state_dict
from rpo_b200 import synth
from tests.common import state_dict

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_bit_exact_vs_reference(prec):
    sd = state_dict("ViT-B/16")
    K, B = 3, 2
    names = ["class 3", "class 41", "class 700"]
    ref = rh.build_reference_customclip(sd, names, K, prec, seed_prompts=11)
    img = synth.make_images(B, 224, seed=5)
    lab = synth.make_labels(B, len(names))
    loss, gt, gi = rh.reference_step(ref, img, lab)
    logits = rh.reference_logits(ref, img)
    tokens = rh.tokenize([f"a photo of a {n}." for n in names])
    om = OracleModel(convert_state_dict(sd, prec), tokens, K, prec)
    tp = ref.prompt_learner.text_prompt.detach()
    ip = ref.prompt_learner.img_prompt.detach()
    l2, gt2, gi2 = om.step(img, tp, ip, lab)
    assert loss.item() == l2.item()
    assert torch.equal(gt, gt2) and torch.equal(gi, gi2)
    assert torch.equal(logits, om.logits(img, tp, ip))
    assert torch.equal(ref.len_prompts, om.len_prompts)
