"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import functools
import os

import numpy as np
import torch

from rpo_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["cfg1_fp32", "cfg1_fp16", "k24_ragged_fp32", "k24_ragged_fp16",
                # edges: len_prompts + K == 77 next to short prompts with one image; a single prompt pair
                "edge_maxlen_fp32", "edge_maxlen_fp16", "edge_k1_fp32", "edge_k1_fp16"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=1)
def tokens_table():
    z = np.load(os.path.join(GOLDEN, "tokens_class1000.npz"))
    T = int(z["context_length"])
    t = z["tokens"]
    full = np.zeros((t.shape[0], T), np.int64)
    full[:, :t.shape[1]] = t
    return torch.from_numpy(full)


def class_tokens(class_ids):
    return tokens_table()[torch.as_tensor(list(class_ids), dtype=torch.int64)]


@functools.lru_cache(maxsize=4)
def state_dict(arch_name, seed=0):
    return synth.make_state_dict(synth.ARCHS[arch_name], seed=seed)


def rel_err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


# Gradients are judged against the TRUTH: the same function evaluated in float64 on the weight / prompt VALUES the
# 16-bit (or fp32) model holds (oracle "fp64" mode).  Three numbers per tensor, all printed, relative to the
# tensor's max magnitude:
#   e_ours = |ours - truth|,  e_ref = |reference in the same precision - truth|,  direct = |ours - reference|.
# fp32: e_ours <= 1e-5 (north_star).
# 16-bit: measured on B200 (gpurun_out/r2a_pytest.log, 60 tensors): the reference's OWN fp16 gradients sit 1.7e-3 ..
# 4.3e-3 from the truth on small shapes (every activation tensor of the chain is rounded to 11 bits), 1e-2 .. 8e-2 at
# batch 32, and 0.3 .. 1.1 (!) for the text prompt at 1000 classes: gradients of 1e-6 .. 1e-3 fall into fp16's
# subnormal range in torch's backward.  This path rounds the forward at the same points but runs the backward on
# gradients scaled by 2^12 with f32 accumulation: 1.7e-3 .. 4e-3 on the shapes where the reference is that good,
# <= 2.1e-2 everywhere, 2.0e-3 .. 2.9e-3 at the 1000-class shapes.  The bar is therefore "at least as close to the
# truth as the reference's own run in that precision, or at the precision's floor": e_ours <= max(1.1 e_ref, FLOOR)
# (1.1: both errors are draws of the same rounding noise) -- and wherever the reference itself is within DIRECT / 2
# of the truth the direct distance must be below DIRECT (fp16: 5e-3).
GRAD_TRUTH_TOL = {"fp32": 1e-5}
FLOOR = {"fp16": 4e-3, "bf16": 3.2e-2}
DIRECT = {"fp32": 2e-5, "fp16": 5e-3, "bf16": 4e-2}


def check_grads(prec, ours, ref_same_prec, truth, what):
    e_ours, e_ref, direct = rel_err(ours, truth), rel_err(ref_same_prec, truth), rel_err(ours, ref_same_prec)
    print(f"    grad {what:5s} [{prec}]: e_ours={e_ours:.3e} e_ref={e_ref:.3e} direct={direct:.3e}")
    if prec == "fp32":
        assert e_ours <= GRAD_TRUTH_TOL[prec], f"{what}: {e_ours:.3e} from the float64 truth"
        assert direct <= DIRECT[prec], f"{what}: {direct:.3e} from the fp32 reference"
        return
    assert e_ours <= max(1.1 * e_ref, FLOOR[prec]), \
        f"{what}: {e_ours:.3e} from the truth; the reference's own {prec} run is {e_ref:.3e} away"
    assert direct <= max(DIRECT[prec], e_ours + e_ref), f"{what}: {direct:.3e} vs the {prec} reference"
    if e_ref <= 0.5 * DIRECT[prec]:
        assert direct <= DIRECT[prec], f"{what}: {direct:.3e} vs the {prec} reference (itself {e_ref:.3e} from the truth)"


def truth_grads(sd, prec, tokens, K, image, tp, ip, label, big=False):
    from oracle.rpo_oracle import OracleModel, convert_state_dict
    """Prompt gradients of the float64 truth model (fp32 on the GPU for the 1000-class shapes, whose float64
    autograd graph does not fit; fp32 sits ~2e-6 from float64, far below the 16-bit bars)."""
    tprec = "fp32" if big else "fp64"
    vals = convert_state_dict(sd, "fp64", prec) if tprec == "fp64" else \
        {k: (v.float() if v.is_floating_point() else v) for k, v in convert_state_dict(sd, prec).items()}
    om = OracleModel(vals, tokens, K, tprec, device="cuda:0")
    _, gt, gi = om.step(image, tp.to(om.dtype), ip.to(om.dtype), label)
    gt, gi = gt.double().cpu(), gi.double().cpu()
    del om
    torch.cuda.empty_cache()
    return gt, gi
