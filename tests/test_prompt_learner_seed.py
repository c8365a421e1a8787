"""PromptLearner initialisation (trainers/rpo.py:60-88) draws its noise from the GLOBAL torch RNG: text noise first,
then visual noise.  A seeded run of rpo_b200.model.PromptLearner must therefore start from bit-identical prompts to the
reference's.  The golden files hold the prompts the UNMODIFIED reference drew under torch.manual_seed(0)
(oracle/ref_harness.py::build_reference_customclip, oracle/make_golden.py)."""
from types import SimpleNamespace

import pytest
import torch

from rpo_b200 import synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP, PromptLearner
from tests.common import GOLDEN_CASES, load_golden, state_dict


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_seeded_prompt_learner_reproduces_the_reference_prompts(name):
    g = load_golden(name)
    prec = name.rsplit("_", 1)[1]
    K = int(g["K"])
    arch = synth.ARCHS["ViT-B/16"]
    clip = SyntheticCLIP(state_dict("ViT-B/16", 0), prec)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    torch.manual_seed(0)
    pl = PromptLearner(cfg, clip)
    assert pl.text_prompt.dtype == clip.dtype and pl.img_prompt.dtype == clip.dtype
    assert torch.equal(pl.text_prompt.detach().float(), torch.from_numpy(g["text_prompt"]))
    assert torch.equal(pl.img_prompt.detach().float(), torch.from_numpy(g["img_prompt"]))
    assert [n for n, _ in pl.named_parameters()] == ["text_prompt", "img_prompt"]  # checkpoint format


def test_custom_clip_draws_the_prompts_first():
    """CustomCLIP.__init__ must not consume the global RNG before the PromptLearner does (trainers/rpo.py:99-101)."""
    g = load_golden("cfg1_fp32")
    K = int(g["K"])
    arch = synth.ARCHS["ViT-B/16"]
    tokens = torch.from_numpy(g["tokens"].astype("int64"))
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC="fp32")),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    torch.manual_seed(0)
    model = CustomCLIP(cfg, [f"c{i}" for i in range(tokens.shape[0])], "a photo of a _.",
                       SyntheticCLIP(state_dict("ViT-B/16", 0), "fp32"), tokens=tokens)
    assert torch.equal(model.prompt_learner.text_prompt.detach().float(), torch.from_numpy(g["text_prompt"]))
    assert torch.equal(model.prompt_learner.img_prompt.detach().float(), torch.from_numpy(g["img_prompt"]))
