"""Pipelined step (rpo_b200.runner.StepRunner(pipeline=True)) against the plain one on one GPU, over the SM budget of
the context pass and the stream-priority switch.  ViT-B/16, K=24, batch 32, fp16; CUDA-event time of `--steps` steps
after warm-up, inputs rotating through 8 batches."""
import argparse
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import synthetic_tokens
from rpo_b200 import synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from rpo_b200.runner import StepRunner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--cls", type=int, default=100)
    ap.add_argument("--sms", type=int, nargs="+", default=[0, 132, 120, 108, 96, 84])
    ap.add_argument("--prio", type=int, nargs="+", default=[1, 0])
    a = ap.parse_args()
    arch = synth.ARCHS["ViT-B/16"]
    K, C, B, prec = 24, a.cls, 32, "fp16"
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(arch, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    pool = [synth.make_images(B, arch.image_resolution, seed=1234 + i).to(dev) for i in range(8)]
    labels = [((torch.arange(B) + i) % C).to(torch.int64).to(dev) for i in range(8)]

    def fresh():
        torch.manual_seed(0)
        m = CustomCLIP(cfg, synth.synthetic_classnames(C), "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=synthetic_tokens(C), max_batch=B).to(dev)
        m.prompt_learner.train()
        return m

    def time_runner(r):
        r.image.copy_(pool[0])
        r.label.copy_(labels[0])
        r.prepare(warmup=2)

        def step(i):
            r.image.copy_(pool[i % 8], non_blocking=True)
            r.label.copy_(labels[i % 8], non_blocking=True)
            r.step()

        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps, float(r.loss.item())

    m = fresh()
    ms, loss = time_runner(StepRunner(m, B))
    print(f"plain                       : {ms:7.3f} ms/step {B / ms * 1e3:8.0f} img/s loss {loss:.4f}", flush=True)
    del m
    m = fresh()
    for prio in a.prio:
        os.environ["RPO_PIPE_PRIO"] = str(prio)
        for sms in a.sms:
            ms, loss = time_runner(StepRunner(m, B, pipeline=True, context_sms=sms))
            print(f"pipeline prio={prio} ctx_sms={sms:3d}: {ms:7.3f} ms/step {B / ms * 1e3:8.0f} img/s loss {loss:.4f}",
                  flush=True)


if __name__ == "__main__":
    main()
