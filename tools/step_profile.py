"""Per-launch device times of one training step (BASELINE config 2 by default) measured with CUDA
events inside a back-to-back, warmed-up run (librpo_b200's launch profiler), aggregated by kernel
site + shape.  Unlike an ncu launch list these are at the clocks of the real step.

    python tools/step_profile.py [--ncls 100] [--batch 32] [--K 24] [--out profiles/x.txt]
"""
import argparse, collections, ctypes as C, os, statistics, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpo_b200 import _lib, synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from rpo_b200.runner import StepRunner
from bench import synthetic_tokens


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="ViT-B/16")
    ap.add_argument("--ncls", type=int, default=100)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--K", type=int, default=24)
    ap.add_argument("--prec", default="fp16")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    arch = synth.ARCHS[a.arch]
    sd = synth.make_state_dict(arch, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=a.K, PREC=a.prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    dev = torch.device("cuda:0")
    model = CustomCLIP(cfg, synth.synthetic_classnames(a.ncls), "a photo of a _.", SyntheticCLIP(sd, a.prec),
                       tokens=synthetic_tokens(a.ncls), max_batch=a.batch).to(dev)
    model.prompt_learner.train()
    runner = StepRunner(model, a.batch, use_graph=False)
    runner.image.copy_(synth.make_images(a.batch, arch.image_resolution).to(dev))
    runner.label.copy_(synth.make_labels(a.batch, a.ncls).to(dev))
    runner.prepare(warmup=5)
    lib = _lib.load()
    st = _lib.stream_ptr(dev)
    buf = C.create_string_buffer(1 << 20)
    per = collections.OrderedDict()
    totals = []
    for rep in range(a.reps):
        for _ in range(3):
            runner.step()  # keep the clocks up
        _lib.check(lib.rpo_profile_begin(st))
        runner.step()
        n = lib.rpo_profile_end(buf, len(buf))
        rows = [l.split("\t") for l in buf.value.decode().splitlines()]
        totals.append(sum(float(r[1]) for r in rows))
        for i, r in enumerate(rows):
            per.setdefault((i, r[0], r[2] if len(r) > 2 else ""), []).append(float(r[1]))
    agg = collections.OrderedDict()
    for (i, where, tag), v in per.items():
        d = agg.setdefault((where, tag), [0, 0.0])
        d[0] += 1
        d[1] += statistics.median(v)
    total = sum(v for _, v in agg.values())
    lines = [f"{a.arch} K={a.K} C={a.ncls} B={a.batch} {a.prec}: one step = {len(per)} launches, "
             f"{statistics.median(totals) / 1e3:.3f} ms (CUDA events, back-to-back, median of {a.reps})",
             f"{'site':24s} {'what':58s} {'n':>4s} {'total us':>9s} {'avg us':>8s} {'share':>6s}"]
    for (where, tag), (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{where:24s} {tag[:58]:58s} {c:4d} {v:9.1f} {v / c:8.2f} {100 * v / total:5.1f}%")
    text = "\n".join(lines)
    print(text)
    if a.out:
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
