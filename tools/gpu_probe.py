"""Prints an error table for the CUDA path vs goldens / oracle (diagnostics, no asserts)."""
import sys, os, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle.rpo_oracle import OracleModel, convert_state_dict
from rpo_b200 import _lib, synth
from tests.common import GOLDEN_CASES, class_tokens, load_golden, rel_err, state_dict
from tests.test_gpu_e2e import build_model, set_prompts, step, eval_logits

def golden():
    for name in GOLDEN_CASES:
        g = load_golden(name); prec = name.rsplit("_", 1)[1]
        g32 = load_golden(name.replace("fp16", "fp32"))
        K, B = int(g["K"]), int(g["B"])
        tokens = torch.from_numpy(g["tokens"].astype(np.int64))
        for backend in (_lib.GEMM_AUTO, _lib.GEMM_SIMT):
            model, arch, _ = build_model("ViT-B/16", prec, K, tokens, backend)
            set_prompts(model, torch.from_numpy(g["text_prompt"]), torch.from_numpy(g["img_prompt"]))
            image = synth.make_images(B, arch.image_resolution).cuda(); label = synth.make_labels(B, tokens.shape[0]).cuda()
            loss, gt, gi = step(model, image, label); logits = eval_logits(model, image)
            T = torch.from_numpy
            print(f"{name} backend={backend}: dloss={abs(loss.item()-float(g['loss'])):.2e} dlogits={(logits-T(g['logits'])).abs().max():.2e} "
                  f"gt={rel_err(gt,T(g['grad_text_prompt'])):.2e} gi={rel_err(gi,T(g['grad_img_prompt'])):.2e} | vs fp32 golden: "
                  f"gt={rel_err(gt,T(g32['grad_text_prompt'])):.2e} gi={rel_err(gi,T(g32['grad_img_prompt'])):.2e} "
                  f"dlogits={(logits-T(g32['logits'])).abs().max():.2e} dloss={abs(loss.item()-float(g32['loss'])):.2e}")
            eng = model._engine; S = arch.n_patch + 1; lp = model.len_prompts; Mc_t = int(lp.sum())
            ev, et = [], []
            for layer in range(arch.vision_layers):
                x = eng.debug_fetch(0, layer).float().cpu()
                got = torch.stack([x[r] if r < S else x[B*S + (r-S)] for r in g["rows_v"].tolist()])
                ev.append(rel_err(got, T(g["taps_v"][layer])))
            for layer in range(arch.transformer_layers):
                x = eng.debug_fetch(1, layer).float().cpu()
                got = torch.stack([x[r] if r < int(lp[0]) else x[Mc_t + (r-int(lp[0]))] for r in g["rows_t"].tolist()])
                et.append(rel_err(got, T(g["taps_t"][layer])))
            print("   taps_v", " ".join(f"{e:.1e}" for e in ev)); print("   taps_t", " ".join(f"{e:.1e}" for e in et))
            del model

def fp64_check():
    # tiny fp32: who is closer to fp64 truth, the CUDA path or torch fp32 on the GPU?
    for arch_name, K, ids, B in (("tiny", 5, [3, 77, 512], 2), ("small", 8, [0, 10, 100, 999], 5)):
        tokens = class_tokens(ids)
        model, arch, sd = build_model(arch_name, "fp32", K, tokens)
        tp, ip = synth.make_prompt_init(sd, K); set_prompts(model, tp, ip)
        image = synth.make_images(B, arch.image_resolution); label = synth.make_labels(B, len(ids))
        loss, gt, gi = step(model, image.cuda(), label.cuda())
        sd32 = convert_state_dict(sd, "fp32")
        o_gpu = OracleModel(sd32, tokens, K, "fp32", device="cuda:0")
        l_g, gt_g, gi_g = o_gpu.step(image, tp, ip, label)
        o_cpu = OracleModel(sd32, tokens, K, "fp32", device="cpu")
        l_c, gt_c, gi_c = o_cpu.step(image, tp, ip, label)
        # fp64 truth
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd32.items()}
        import oracle.rpo_oracle as ro
        ro.PREC_DTYPE["fp64"] = torch.float64
        old_ln = ro.layer_norm
        ro.layer_norm = lambda x, w, b: torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
        o64 = OracleModel(sd64, tokens, K, "fp64", device="cpu")
        l64, gt64, gi64 = o64.step(image.double(), tp.double(), ip.double(), label)
        ro.layer_norm = old_ln
        print(f"{arch_name} fp32 vs fp64 truth: loss cuda {abs(loss.item()-l64.item()):.2e} torch-gpu {abs(l_g.item()-l64.item()):.2e} torch-cpu {abs(l_c.item()-l64.item()):.2e}")
        print(f"    gt: cuda {rel_err(gt, gt64):.2e} torch-gpu {rel_err(gt_g, gt64):.2e} torch-cpu {rel_err(gt_c, gt64):.2e};  gi: cuda {rel_err(gi, gi64):.2e} torch-gpu {rel_err(gi_g, gi64):.2e} torch-cpu {rel_err(gi_c, gi64):.2e}")

if __name__ == "__main__":
    golden(); fp64_check()
