"""Class-sharded text tower against plain data parallelism, same process group, same box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/shard_bench.py [--steps 30] [--cls 100 1000]

For every class count: ViT-B/16, K=24, batch 32 per GPU, fp16 (BASELINE.json configs[1] / configs[3] geometry), one
StepRunner per mode (plain data parallelism, `model.shard_text()`, pipelined context rows, both); CUDA-event time of `steps` steps after warm-up, max over ranks.
Rank 0 prints one JSON line per run."""
import argparse
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from bench import synthetic_tokens
from rpo_b200 import synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from rpo_b200.runner import StepRunner


def run(C, shard, pipeline, steps, warmup, B=32, K=24, arch_name="ViT-B/16", prec="fp16"):
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    arch = synth.ARCHS[arch_name]
    sd = synth.make_state_dict(arch, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    torch.manual_seed(0)
    model = CustomCLIP(cfg, synth.synthetic_classnames(C), "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=synthetic_tokens(C), max_batch=B).to(dev)
    model.prompt_learner.train()
    if shard:
        model.shard_text()
    r = StepRunner(model, B, use_graph=True, process_group=dist.group.WORLD, world_size=world, pipeline=pipeline)
    pool = [synth.make_images(B, arch.image_resolution, seed=1234 + 97 * rank + i).to(dev) for i in range(8)]
    labels = [((torch.arange(B) + i + rank) % C).to(torch.int64).to(dev) for i in range(8)]
    r.image.copy_(pool[0])
    r.label.copy_(labels[0])
    r.prepare(warmup=3)

    def step(i):
        r.image.copy_(pool[i % 8], non_blocking=True)
        r.label.copy_(labels[i % 8], non_blocking=True)
        r.step()

    for i in range(warmup):
        step(i)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    losses = torch.tensor([float(r.loss.item())], dtype=torch.float64, device=dev)
    dist.all_reduce(losses)
    if rank == 0:
        ms = float(t[0])
        print(json.dumps({"n_gpus": world, "n_cls": C, "K": K, "batch_per_gpu": B, "shard_text": bool(shard), "pipeline": bool(pipeline),
                          "ms_per_step": ms, "images_per_s": B * world / ms * 1e3,
                          "mean_loss_after": float(losses[0]) / world, "launches_per_step": r.launches_per_step,
                          "device_bytes": r.eng.device_bytes(),
                          "prompt_checksum": float(model.prompt_learner.text_prompt.data.float().abs().sum())}),
              flush=True)
    del r, model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cls", type=int, nargs="+", default=[100, 1000])
    ap.add_argument("--modes", nargs="+", default=["plain", "shard", "pipe", "shard+pipe"])
    a = ap.parse_args()
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    for C in a.cls:
        for mode in a.modes:
            run(C, "shard" in mode, "pipe" in mode, a.steps, a.warmup)
    dist.barrier()
    dist.destroy_process_group()
