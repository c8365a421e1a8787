"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):   python oracle/make_golden.py

For every case it builds the synthetic weights (rpo_b200.synth, seeded), loads them into the
reference's own clip.model.build_model + trainers/rpo.py::CustomCLIP (oracle/ref_harness.py), runs one
forward+CE+backward and one eval forward, and stores: the token ids produced by the reference
tokenizer, the prompt values drawn by the reference PromptLearner, loss, eval logits, both prompt
gradients and a few residual-stream rows after every block (forward hooks on resblocks).
Weights and images are NOT stored: they are regenerated from their seeds on the test machine.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from rpo_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # BASELINE.json configs[0]
    "cfg1_fp32": dict(arch="ViT-B/16", K=4, class_ids=[0, 1], B=2, prec="fp32"),
    "cfg1_fp16": dict(arch="ViT-B/16", K=4, class_ids=[0, 1], B=2, prec="fp16"),
    # K=24 with ragged prompt lengths (1/2/3-digit class ids tokenise to 9/10/11 tokens)
    "k24_ragged_fp32": dict(arch="ViT-B/16", K=24, class_ids=[0, 5, 17, 123, 999, 42, 7, 256, 1, 64], B=3, prec="fp32"),
    "k24_ragged_fp16": dict(arch="ViT-B/16", K=24, class_ids=[0, 5, 17, 123, 999, 42, 7, 256, 1, 64], B=3, prec="fp16"),
    # edges of the index arithmetic (trainers/rpo.py:137,149,177): a class prompt whose K prompt slots end exactly
    # at position 76 (len_prompts + K == 77) next to a 9- and an 8-token prompt, one image; and a single prompt pair
    "edge_maxlen_fp32": dict(arch="ViT-B/16", K=4, names=["class 7", " ".join(["dog"] * 66), "x"], B=1, prec="fp32"),
    "edge_maxlen_fp16": dict(arch="ViT-B/16", K=4, names=["class 7", " ".join(["dog"] * 66), "x"], B=1, prec="fp16"),
    "edge_k1_fp32": dict(arch="ViT-B/16", K=1, class_ids=[0, 10, 100], B=2, prec="fp32"),
    "edge_k1_fp16": dict(arch="ViT-B/16", K=1, class_ids=[0, 10, 100], B=2, prec="fp16"),
}
TAP_ROWS_V = [0, 100]  # cls row and one patch row; the first and last prompt rows are appended
TAP_ROWS_T = [0, 4]


def run_case(name, spec):
    arch = synth.ARCHS[spec["arch"]]
    sd = synth.make_state_dict(arch, seed=0)
    names = spec.get("names") or [f"class {i}" for i in spec["class_ids"]]
    K, B, prec = spec["K"], spec["B"], spec["prec"]
    C = len(names)
    model = rh.build_reference_customclip(sd, names, K, prec, seed_prompts=0)
    image = synth.make_images(B, arch.image_resolution, seed=1234)
    label = synth.make_labels(B, C)

    vt, tt = [], []
    hooks = []
    for blk in model.img_transformer.resblocks:
        hooks.append(blk.register_forward_hook(lambda m, i, o: vt.append(o.detach().float())))
    for blk in model.text_transformers.resblocks:
        hooks.append(blk.register_forward_hook(lambda m, i, o: tt.append(o.detach().float())))
    loss, g_text, g_img = rh.reference_step(model, image, label)
    for h in hooks:
        h.remove()
    logits = rh.reference_logits(model, image)

    S = arch.n_patch + 1
    lp = model.len_prompts
    rows_v = TAP_ROWS_V + [S, S + K - 1]
    # vision taps: [layers, len(rows_v), D] for image 0 (hook output is [L, N, D])
    taps_v = torch.stack([o[rows_v, 0, :] for o in vt])
    # text taps for class 0: two context rows + first/last prompt rows
    rows_t = TAP_ROWS_T + [int(lp[0]), int(lp[0]) + K - 1]
    taps_t = torch.stack([o[rows_t, 0, :] for o in tt])
    out = dict(
        tokens=model.text_tokenized.numpy().astype(np.int32),
        class_ids=np.asarray(spec.get("class_ids", []), np.int32),  # empty: free-form class names
        K=np.int32(K), B=np.int32(B),
        text_prompt=model.prompt_learner.text_prompt.detach().float().numpy(),
        img_prompt=model.prompt_learner.img_prompt.detach().float().numpy(),
        loss=np.float32(loss.item()),
        logits=logits.float().numpy(),
        grad_text_prompt=g_text.float().numpy(),
        grad_img_prompt=g_img.float().numpy(),
        rows_v=np.asarray(rows_v, np.int32), taps_v=taps_v.numpy(),
        rows_t=np.asarray(rows_t, np.int32), taps_t=taps_t.numpy(),
    )
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: len_prompts={lp.tolist()} loss={loss.item():.6f} |g_text|max={g_text.abs().max():.3e} "
          f"|g_img|max={g_img.abs().max():.3e} -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


def tokens_table():
    """Token ids of 'a photo of a class {i}.' for i < 1000 from the reference tokenizer
    (clip/clip.py:185-221), so the GPU box can build C<=1000 synthetic classes without the BPE vocab."""
    toks = rh.tokenize([f"a photo of a class {i}." for i in range(1000)]).numpy().astype(np.int32)
    used = int((toks != 0).any(axis=0).nonzero()[0].max()) + 1
    assert (toks[:, used:] == 0).all()
    path = os.path.join(GOLDEN_DIR, "tokens_class1000.npz")
    np.savez_compressed(path, tokens=toks[:, :used], context_length=np.int32(toks.shape[1]))
    print("tokens:", toks.shape, "non-zero columns:", used, "->", path)


if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    if not only:
        tokens_table()
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_case(name, spec)
