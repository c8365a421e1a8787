"""Device time of the forward and of the backward+update halves of one step (CUDA events, no graph), with the
text tower on the side stream (default) and serialised on one stream (RPO_SINGLE_STREAM=1 in the environment)."""
import os, statistics, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rpo_b200 import synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from rpo_b200.runner import StepRunner
from bench import synthetic_tokens


def main():
    arch = synth.ARCHS["ViT-B/16"]
    K, C, B, prec = 24, 100, 32, "fp16"
    sd = synth.make_state_dict(arch, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    dev = torch.device("cuda:0")
    model = CustomCLIP(cfg, synth.synthetic_classnames(C), "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=synthetic_tokens(C), max_batch=B).to(dev)
    model.prompt_learner.train()
    r = StepRunner(model, B, use_graph=False)
    r.image.copy_(synth.make_images(B, arch.image_resolution).to(dev))
    r.label.copy_(synth.make_labels(B, C).to(dev))
    r.prepare(warmup=5)
    pl = model.prompt_learner
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    f, b = [], []
    for it in range(30):
        ev[0].record()
        r.eng.forward(r.image, pl.text_prompt.data, pl.img_prompt.data, r.label)
        ev[1].record()
        r.eng.backward()
        r._update()
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 5:
            f.append(ev[0].elapsed_time(ev[1]))
            b.append(ev[1].elapsed_time(ev[2]))
    print(f"single_stream={os.environ.get('RPO_SINGLE_STREAM', '0')}  forward {statistics.median(f):.3f} ms  "
          f"backward+update {statistics.median(b):.3f} ms  total {statistics.median(f) + statistics.median(b):.3f} ms")


if __name__ == "__main__":
    main()
