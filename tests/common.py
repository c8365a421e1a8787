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
