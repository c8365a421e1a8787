"""The thing bench.py times -- runner.StepRunner: CUDA-graph step with the fused SGD kernel -- against the oracle
driven the way the reference trainer drives the model (trainers/rpo.py:290-316): loss = model(image, label);
optim.zero_grad(); loss.backward(); optim.step() with torch.optim.SGD(momentum 0.9, weight decay 5e-4; Dassl's
OPTIM defaults), a different batch every step, over several steps, from the same initial state.

fp32: same trajectory to 1e-5.  fp16: the reference's SGD runs in fp16 (parameter, gradient and momentum buffer are
fp16 tensors), this path keeps the gradient and the momentum in f32 and rounds once per step when it stores the
parameter, so the two trajectories agree to a few fp16 ulps of the parameters (an update of lr * g ~ 1e-5 is below
one ulp of most entries: both sides quantise, slightly differently); a float64 trajectory with the same
hyper-parameters sits between them and is the reference point for the bound.
"""
from types import SimpleNamespace

import pytest
import torch

from oracle.rpo_oracle import OracleModel, convert_state_dict
from rpo_b200 import synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP
from rpo_b200.runner import StepRunner
from tests.common import class_tokens, rel_err, state_dict

pytestmark = pytest.mark.gpu

LR, MOM, WD = 0.01, 0.9, 5e-4  # configs/trainers/RPO/main_K24.yaml:15-16 + Dassl defaults


def make_model(arch_name, prec, K, tokens):
    arch = synth.ARCHS[arch_name]
    sd = state_dict(arch_name, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    model = CustomCLIP(cfg, [f"c{i}" for i in range(tokens.shape[0])], "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=tokens).to("cuda:0")
    tp, ip = synth.make_prompt_init(sd, K)
    with torch.no_grad():
        model.prompt_learner.text_prompt.copy_(tp.to(model.dtype))
        model.prompt_learner.img_prompt.copy_(ip.to(model.dtype))
    return model, arch, sd


def oracle_trajectory(sd, prec, values_of, tokens, K, tp0, ip0, batches, lr):
    om = OracleModel(convert_state_dict(sd, prec, values_of) if prec == "fp64" else convert_state_dict(sd, prec),
                     tokens, K, prec, device="cuda:0")
    tp = tp0.detach().to("cuda:0", om.dtype).clone().requires_grad_(True)
    ip = ip0.detach().to("cuda:0", om.dtype).clone().requires_grad_(True)
    opt = torch.optim.SGD([tp, ip], lr=lr, momentum=MOM, weight_decay=WD)
    losses = []
    for image, label in batches:
        loss = om.forward(image.cuda(), tp, ip, label.cuda(), training=True)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.item()))
    del om
    torch.cuda.empty_cache()
    return losses, tp.detach().double().cpu(), ip.detach().double().cpu()


@pytest.mark.parametrize("arch_name,prec,K,class_ids,B,lr", [
    ("small", "fp32", 8, [0, 10, 100, 999], 5, LR),
    ("small", "fp16", 8, [0, 10, 100, 999], 5, LR),
    ("ViT-B/16", "fp16", 24, list(range(0, 1000, 53)), 8, LR),
    ("small", "fp16", 8, [0, 10, 100, 999], 5, 1.0),   # large steps: every update is far above one ulp
])
@pytest.mark.parametrize("use_graph", [True, False], ids=["graph", "eager"])
def test_step_runner_follows_the_reference_trainer(arch_name, prec, K, class_ids, B, lr, use_graph):
    if arch_name == "ViT-B/16" and not use_graph:
        pytest.skip("one big case is enough")
    steps = 4
    tokens = class_tokens(class_ids)
    model, arch, sd = make_model(arch_name, prec, K, tokens)
    tp0 = model.prompt_learner.text_prompt.detach().clone()
    ip0 = model.prompt_learner.img_prompt.detach().clone()
    batches = [(synth.make_images(B, arch.image_resolution, seed=100 + i),
                (synth.make_labels(B, len(class_ids)) + i) % len(class_ids)) for i in range(steps)]
    runner = StepRunner(model, B, lr=lr, momentum=MOM, weight_decay=WD, use_graph=use_graph)
    runner.image.copy_(batches[0][0])
    runner.label.copy_(batches[0][1])
    runner.prepare(warmup=3)
    # the warm-up steps must not leak into the trajectory (prompts, momentum, first-step flag)
    assert torch.equal(model.prompt_learner.text_prompt.data, tp0)
    assert torch.equal(model.prompt_learner.img_prompt.data, ip0)
    assert float(runner.mom_buf.abs().max()) == 0.0 and int(runner.first.item()) == 1
    losses = []
    for image, label in batches:
        runner.image.copy_(image)
        runner.label.copy_(label)
        runner.step()
        losses.append(runner.loss.clone())
    torch.cuda.synchronize()
    losses = [float(x) for x in losses]
    tp1 = model.prompt_learner.text_prompt.detach().double().cpu()
    ip1 = model.prompt_learner.img_prompt.detach().double().cpu()
    ref_l, ref_tp, ref_ip = oracle_trajectory(sd, prec, None, tokens, K, tp0, ip0, batches, lr)
    tru_l, tru_tp, tru_ip = oracle_trajectory(sd, "fp64", prec, tokens, K, tp0, ip0, batches, lr)
    # how far the prompts moved: errors are relative to the size of the total update, not of the prompts
    move_t = (tru_tp - tp0.double().cpu()).abs().max().item()
    move_i = (tru_ip - ip0.double().cpu()).abs().max().item()

    def upd_err(a, b, move):
        return (a - b).abs().max().item() / move

    e_ours = (upd_err(tp1, tru_tp, move_t), upd_err(ip1, tru_ip, move_i))
    e_ref = (upd_err(ref_tp, tru_tp, move_t), upd_err(ref_ip, tru_ip, move_i))
    dl = max(abs(a - b) for a, b in zip(losses, ref_l))
    dl_truth = max(abs(a - b) for a, b in zip(losses, tru_l))
    print(f"{arch_name}/{prec} lr={lr} graph={use_graph}: losses ours {losses} ref {ref_l} truth {tru_l}; "
          f"max |dloss| vs ref {dl:.3e} vs truth {dl_truth:.3e}; total update max |text| {move_t:.3e} |img| {move_i:.3e}; "
          f"update error / update: ours {e_ours[0]:.3e} {e_ours[1]:.3e}  reference {e_ref[0]:.3e} {e_ref[1]:.3e}")
    tol = {"fp32": 1e-5, "fp16": 1e-3}[prec]
    for a, b in zip(losses, ref_l):
        assert abs(a - b) <= tol * max(1.0, abs(b)) * (4 if lr >= 1.0 else 1)
    if prec == "fp32":
        assert e_ours[0] <= 1e-4 and e_ours[1] <= 1e-4
        assert rel_err(tp1, ref_tp) <= 1e-6 and rel_err(ip1, ref_ip) <= 1e-6
    else:
        # at least as close to the float64 trajectory as the reference's fp16 optimiser, up to one storage rounding
        ulp_t = float(tp0.abs().max()) * 2 ** -10 / move_t
        ulp_i = float(ip0.abs().max()) * 2 ** -10 / move_i
        assert e_ours[0] <= max(e_ref[0], 0.02) + ulp_t, (e_ours, e_ref, ulp_t)
        assert e_ours[1] <= max(e_ref[1], 0.02) + ulp_i, (e_ours, e_ref, ulp_i)


# ---- the drop-in trainer: rpo_b200.trainer.RPO.forward_backward (fast path) ----------------------------------------
def _stub_trainer(model, lr):
    """rpo_b200.trainer.RPO over `object` (Dassl is not installed here) with the attributes TrainerX provides."""
    from rpo_b200 import trainer
    t = trainer.RPO.__new__(trainer.RPO)
    t.model = model
    t.device = model.w_mm.device
    t.optim = torch.optim.SGD(model.prompt_learner.parameters(), lr=lr, momentum=MOM, weight_decay=WD)
    t.sched = torch.optim.lr_scheduler.StepLR(t.optim, step_size=1, gamma=0.5)
    t.scaler = None
    t.batch_idx, t.num_batches = 0, 2
    t.update_lr = lambda: t.sched.step()
    return t


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_trainer_fast_path_follows_the_autograd_path(prec, monkeypatch):
    """forward_backward through the uploader + CUDA-graph step (default) against the reference-shaped path (autograd
    Function + torch.optim.SGD.step) over two 'epochs' of two batches with an lr schedule in between, from host
    batches as a DataLoader hands them over (pageable float32)."""
    tokens = class_tokens([0, 10, 100, 999])
    K, B, steps = 8, 5, 4
    outs = []
    for fast in ("1", "0"):
        monkeypatch.setenv("RPO_B200_FAST", fast)
        model, arch, sd = make_model("small", prec, K, tokens)
        model.prompt_learner.train()
        t = _stub_trainer(model, lr=0.02)
        assert t.fast_path_available() == (fast == "1")
        losses = []
        for i in range(steps):
            t.batch_idx = i % t.num_batches
            batch = {"img": synth.make_images(B, arch.image_resolution, seed=200 + i),
                     "label": (synth.make_labels(B, 4) + i) % 4}
            losses.append(t.forward_backward(batch)["loss"])
        torch.cuda.synchronize()
        if fast == "1":
            assert t._fast["runner"].graph is not None
            losses = losses[1:] + [t._fast["loss"].latest(0)]  # step n reports the loss of step n - 1
            t.export_optimizer_state()
        lr_end = t.optim.param_groups[0]["lr"]
        mom = [t.optim.state[p]["momentum_buffer"].detach().float().cpu() for p in model.prompt_learner.parameters()]
        outs.append((losses, model.prompt_learner.text_prompt.detach().float().cpu(),
                     model.prompt_learner.img_prompt.detach().float().cpu(), lr_end, mom))
    (la, tpa, ipa, lra, ma), (lb, tpb, ipb, lrb, mb) = outs
    print(f"{prec}: fast {la} autograd {lb}; prompts text {rel_err(tpa, tpb):.3e} img {rel_err(ipa, ipb):.3e}; "
          f"momentum text {rel_err(ma[0], mb[0]):.3e} img {rel_err(ma[1], mb[1]):.3e}")
    assert lra == lrb == 0.02 * 0.25
    tol = {"fp32": 1e-5, "fp16": 1e-3}[prec]
    for a, b in zip(la, lb):
        assert abs(a - b) <= tol * max(1.0, abs(b))
    ptol = {"fp32": 1e-6, "fp16": 2 ** -9}[prec]   # fp16: two storage roundings of the parameter
    assert rel_err(tpa, tpb) <= ptol and rel_err(ipa, ipb) <= ptol
    mtol = {"fp32": 1e-4, "fp16": 2e-2}[prec]
    assert rel_err(ma[0], mb[0]) <= mtol and rel_err(ma[1], mb[1]) <= mtol
