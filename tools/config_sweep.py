"""Step throughput on the other BASELINE.json configurations (single GPU): ViT-L/14 bf16 (config 3), ViT-B/16 with 1000
classes (config 4) and the K sweep at batch 64 (config 5).  Synthetic weights / images as in bench.py; CUDA-graph step,
CUDA-event timing, 20 timed steps after 5 warm-up steps."""
import os, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpo_b200 import synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from rpo_b200.runner import StepRunner
from bench import synthetic_tokens


def run(arch_name, prec, K, C, B, steps=20):
    arch = synth.ARCHS[arch_name]
    sd = synth.make_state_dict(arch, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = CustomCLIP(cfg, synth.synthetic_classnames(C), "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=synthetic_tokens(C), max_batch=B).to(dev)
    model.prompt_learner.train()
    r = StepRunner(model, B, use_graph=True)
    r.image.copy_(synth.make_images(B, arch.image_resolution).to(dev))
    r.label.copy_(synth.make_labels(B, C).to(dev))
    r.prepare(warmup=3)
    for _ in range(5):
        r.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss = float(r.loss.item())
    print(f"{arch_name:9s} {prec} K={K:2d} C={C:4d} B={B:2d}: {ms:7.3f} ms/step  {B / ms * 1e3:8.0f} img/s  loss {loss:.4f}  "
          f"workspace {r.eng.device_bytes() / 2**30:.2f} GiB", flush=True)
    del r, model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("ViT-B/16", "fp16", 24, 100, 32)    # config 2 (bench.py)
    run("ViT-L/14", "bf16", 24, 100, 16)    # config 3
    run("ViT-B/16", "fp16", 24, 1000, 32)   # config 4 (per GPU)
    for K in (4, 8, 16, 24, 48):            # config 5 (per GPU)
        run("ViT-B/16", "fp16", K, 1000, 64)
