"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (mlvlab/RPO at /root/reference).

Only usable in the authoring container, where /root/reference is mounted; it cannot travel to the
GPU box.  It is used (a) by oracle/make_golden.py to generate tests/golden/*.npz from the reference's
own `trainers/rpo.py::CustomCLIP` and (b) by tests/test_oracle_vs_reference.py to pin the restatement
in oracle/rpo_oracle.py against the reference itself.

The reference needs `ftfy` and `dassl.*` at import time only (trainers/rpo.py:13-16,
clip/simple_tokenizer.py:6); neither is installed and neither is touched by CustomCLIP/PromptLearner
at run time, so both are replaced by inert stub modules.
"""
import os
import sys
import types
from types import SimpleNamespace

import torch

REFERENCE_ROOT = os.environ.get("RPO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "trainers", "rpo.py"))


def _install_stubs():
    sys.dont_write_bytecode = True  # the reference mount is read-only
    if "ftfy" not in sys.modules:
        ftfy = types.ModuleType("ftfy")
        ftfy.fix_text = lambda s: s
        sys.modules["ftfy"] = ftfy
    if "dassl" not in sys.modules:
        dassl = types.ModuleType("dassl")
        engine = types.ModuleType("dassl.engine")

        class _Registry:
            def register(self):
                return lambda cls: cls

        engine.TRAINER_REGISTRY = _Registry()
        engine.TrainerX = object
        metrics = types.ModuleType("dassl.metrics")
        metrics.compute_accuracy = None
        utils = types.ModuleType("dassl.utils")
        utils.load_pretrained_weights = None
        utils.load_checkpoint = None
        optim = types.ModuleType("dassl.optim")
        optim.build_optimizer = None
        optim.build_lr_scheduler = None
        for m in (dassl, engine, metrics, utils, optim):
            sys.modules[m.__name__] = m


_CACHE = {}


def import_reference():
    """Returns (trainers.rpo module, clip.model module, clip.clip module) of the reference."""
    if "mods" in _CACHE:
        return _CACHE["mods"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    # Import under the reference's own top-level names but without letting this repo's
    # `trainers/` drop-in shadow it: temporarily put the reference first and purge after.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "trainers" or k.startswith("trainers.")
             or k == "clip" or k.startswith("clip.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        ref_rpo = importlib.import_module("trainers.rpo")
        ref_model = importlib.import_module("clip.model")
        ref_clip = importlib.import_module("clip.clip")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # keep the reference modules reachable only through our handles
        for k in list(sys.modules):
            if k == "trainers" or k.startswith("trainers.") or k == "clip" or k.startswith("clip."):
                sys.modules.pop(k)
        sys.modules.update(saved)
    _CACHE["mods"] = (ref_rpo, ref_model, ref_clip)
    return _CACHE["mods"]


def make_cfg(K: int, prec: str, backbone: str = "ViT-B/16", prompt: str = "a photo of a _."):
    return SimpleNamespace(
        TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec, CTX_INIT="")),
        INPUT=SimpleNamespace(SIZE=(224, 224)),
        DATASET=SimpleNamespace(PROMPT=prompt),
        MODEL=SimpleNamespace(BACKBONE=SimpleNamespace(NAME=backbone), INIT_WEIGHTS=""),
    )


def build_reference_clip(state_dict, prec: str):
    """clip.model.build_model(state_dict) -> fp16 CLIP exactly as real checkpoints are converted
    (clip/model.py:403-440); `.float()` for fp32 as trainers/rpo.py:247-249 does."""
    _, ref_model, _ = import_reference()
    sd = {k: v.clone() for k, v in state_dict.items()}
    model = ref_model.build_model(sd)
    if prec == "fp32":
        model.float()
    return model


def build_reference_customclip(state_dict, classnames, K: int, prec: str, seed_prompts: int = 0):
    """The reference CustomCLIP with everything but prompt_learner frozen (trainers/rpo.py:258-260).
    The global torch RNG is seeded right before construction because PromptLearner draws its noise
    from it (trainers/rpo.py:65,79)."""
    ref_rpo, _, _ = import_reference()
    clip_model = build_reference_clip(state_dict, prec)
    cfg = make_cfg(K, prec)
    torch.manual_seed(seed_prompts)
    model = ref_rpo.CustomCLIP(cfg, classnames, cfg.DATASET.PROMPT, clip_model)
    for name, p in model.named_parameters():
        if "prompt_learner" not in name:
            p.requires_grad_(False)
    return model


def reference_step(model, image, label):
    """One forward+backward of the unmodified reference; applies the H6 workaround: on
    CPU `self.text_x.to(device)` aliases, so hand the model a fresh clone each step."""
    base = getattr(model, "_text_x_base", None)
    if base is None:
        base = model.text_x.detach().clone()
        model._text_x_base = base
    model.text_x = base.clone()
    model.prompt_learner.train()
    for p in model.prompt_learner.parameters():
        p.grad = None
    loss = model(image, label)
    loss.backward()
    return loss.detach(), model.prompt_learner.text_prompt.grad.detach().clone(), \
        model.prompt_learner.img_prompt.grad.detach().clone()


def reference_logits(model, image):
    base = getattr(model, "_text_x_base", None)
    if base is None:
        base = model.text_x.detach().clone()
        model._text_x_base = base
    model.text_x = base.clone()
    model.prompt_learner.eval()
    with torch.no_grad():
        out = model(image)
    model.prompt_learner.train()
    return out


def tokenize(texts):
    _, _, ref_clip = import_reference()
    return ref_clip.tokenize(texts)
