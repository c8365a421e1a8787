"""GPU parity of the class-sharded text tower (include/rpo_b200.h stage entry points,
rpo_b200/text_shard.py).  The reference runs every class prompt on every GPU
(trainers/rpo.py:173-192); sharding them must not change a result.

Single-device tests (run with the plain `-m gpu` suite): two class-sharded handles play the two
ranks on one GPU, the collectives are done by hand between their native stages, and the outcome is
compared with the unsharded handle over the global batch and with the oracle.  The real two-process
NCCL path is `test_two_process_nccl` (needs 2 GPUs; skipped otherwise)."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch

from oracle.rpo_oracle import OracleModel, convert_state_dict
from rpo_b200 import _lib, synth
from rpo_b200.clip_weights import SyntheticCLIP
from rpo_b200.model import CustomCLIP, Engine
from rpo_b200.runner import StepRunner
from rpo_b200.text_shard import ClassShard
from tests.common import class_tokens, rel_err, state_dict

pytestmark = pytest.mark.gpu


def make_model(arch_name, prec, K, tokens, device="cuda:0"):
    arch = synth.ARCHS[arch_name]
    sd = state_dict(arch_name, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    model = CustomCLIP(cfg, [f"c{i}" for i in range(tokens.shape[0])], "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=tokens).to(device)
    tp, ip = synth.make_prompt_init(sd, K)
    with torch.no_grad():
        model.prompt_learner.text_prompt.copy_(tp.to(model.dtype))
        model.prompt_learner.img_prompt.copy_(ip.to(model.dtype))
    return model, arch, sd


def sharded_engine(model, rank, world, batch):
    shard = ClassShard(model.text_x.shape[0], rank, world)
    return Engine(model.arch, model.K, model.text_x.shape[0], model.dtype, batch, model.w_mm, model.w_f32,
                  model._index, model.text_x, model.len_prompts, model.gemm_backend, shard, None)


@pytest.mark.parametrize("arch_name,prec,K,class_ids,B", [
    ("tiny", "fp32", 5, [3, 77, 512, 9, 40], 3),          # 5 classes over 2 ranks: parts of 3 and 2 (padded)
    ("small", "fp16", 8, [0, 10, 100, 999, 5, 6, 7], 4),
    ("ViT-B/16", "fp16", 24, list(range(0, 1000, 91)), 4),
])
def test_two_shards_on_one_device_match_unsharded(arch_name, prec, K, class_ids, B):
    world = 2
    tokens = class_tokens(class_ids)
    Cn = len(class_ids)
    model, arch, sd = make_model(arch_name, prec, K, tokens)
    image = synth.make_images(world * B, arch.image_resolution).cuda()
    label = synth.make_labels(world * B, Cn).cuda()
    tp, ip = model.prompt_learner.text_prompt.data, model.prompt_learner.img_prompt.data
    # ---- unsharded handle over the global batch
    full = model.engine(world * B)
    loss_full, logits_full = full.forward(image, tp, ip, label, want_logits=True)
    loss_full, logits_full = loss_full.clone(), logits_full.clone()
    grad_full = full.backward().clone()
    text_feat_full = full.debug_fetch(3, 0).clone()
    # ---- two class-sharded handles, B images each
    engs = [sharded_engine(model, r, world, B) for r in range(world)]
    idt = _lib.RPO_F32
    for e in engs:
        e.text_forward(tp)
    for e in engs:  # all-gather by hand
        for o in engs:
            if o is not e:
                x = o.exchange
                e.exchange.text_feat[x.r0:x.r0 + x.nl].copy_(x.text_feat[x.r0:x.r0 + x.nl])
    losses, logits = [], []
    for r, e in enumerate(engs):
        sl = slice(r * B, (r + 1) * B)
        e.image_forward(image[sl].contiguous(), idt, ip)
        e.logits_forward(label[sl].contiguous(), e.logits[:B])
        losses.append(e.loss.clone())
        logits.append(e.logits[:B].clone())
        e.logits_backward()
    total = sum(e.exchange.d_text_feat.float() for e in engs)  # reduce-scatter (sum) by hand
    for e in engs:
        x = e.exchange
        x.d_text_feat[x.r0:x.r0 + x.nl].copy_(total[x.r0:x.r0 + x.nl].to(x.d_text_feat.dtype))
    flats = []
    for e in engs:
        e.text_backward()
        e.image_backward()
        flats.append(e.grad_flat.clone())
    torch.cuda.synchronize()
    flat = sum(flats) / world
    # text features of the local classes are the unsharded ones, bit for bit (same kernels, same rows)
    for e in engs:
        x = e.exchange
        assert torch.equal(x.text_feat[x.r0:x.r0 + x.nl], text_feat_full[x.r0:x.r0 + x.nl])
        assert torch.all(x.text_feat[Cn * K:] == 0) and torch.all(x.d_text_feat[Cn * K:] == 0)
    # logits are per image: bit-identical to the global batch's rows
    assert torch.equal(torch.cat(logits), logits_full)
    tol = 1e-5 if prec == "fp32" else 1e-3
    assert abs((sum(losses) / world).item() - loss_full.item()) <= tol * max(1.0, abs(loss_full.item()))
    nt = full.n_text
    gtol = 3e-5 if prec == "fp32" else 2e-2  # 16-bit: the two paths round partial sums of B vs 2B images differently
    print(f"{arch_name}/{prec}: text {rel_err(flat[:nt], grad_full[:nt]):.3e} image {rel_err(flat[nt:], grad_full[nt:]):.3e}")
    assert rel_err(flat[:nt], grad_full[:nt]) <= gtol
    assert rel_err(flat[nt:], grad_full[nt:]) <= gtol
    # and against the oracle over the global batch
    om = OracleModel(convert_state_dict(sd, "fp32"), tokens, K, "fp32", device="cuda:0")
    _, ogt, ogi = om.step(image, tp.float(), ip.float(), label)
    otol = 3e-5 if prec == "fp32" else 6e-2
    assert rel_err(flat[:nt].view(K, -1), ogt) <= otol
    assert rel_err(flat[nt:].view(K, -1), ogi) <= otol
    # composite calls on a sharded handle are refused: they would skip the exchange
    with pytest.raises(_lib.RpoError):
        _lib.check(engs[0].lib.rpo_backward(engs[0].handle, _lib.ptr(engs[0].grad_flat), _lib.stream_ptr(engs[0].device)))


def test_world1_shard_is_bit_identical_through_public_surface():
    """model.shard_text(0, 1): stage path + graph segments against the single-call path / single graph."""
    tokens = class_tokens([1, 20, 300, 4, 55])
    K, B = 4, 3
    a, arch, _ = make_model("small", "fp16", K, tokens)
    b, _, _ = make_model("small", "fp16", K, tokens)
    b.shard_text(0, 1)
    image = synth.make_images(B, arch.image_resolution).cuda()
    label = synth.make_labels(B, 5).cuda()
    out = []
    for m in (a, b):
        m.prompt_learner.train()
        loss = m(image, label)
        loss.backward()
        m.prompt_learner.eval()
        with torch.no_grad():
            lg1 = m(image)
            lg2 = m(image)  # cached text features
        m.prompt_learner.train()
        out.append((loss.detach().clone(), m.prompt_learner.text_prompt.grad.clone(),
                    m.prompt_learner.img_prompt.grad.clone(), lg1, lg2))
    assert b._engine.exchange is not None and a._engine.exchange is None
    for x, y in zip(*out):
        assert torch.equal(x, y)
    # StepRunner: five graph segments vs one graph
    losses = []
    for m in (a, b):
        r = StepRunner(m, B, lr=0.02).prepare(warmup=2)
        r.image.copy_(image)
        r.label.copy_(label)
        ls = []
        for _ in range(4):
            r.step()
            ls.append(r.loss.clone())
        torch.cuda.synchronize()
        losses.append(torch.stack(ls).cpu())
    assert torch.equal(losses[0], losses[1]), losses
    assert losses[0][-1] < losses[0][0]
    assert torch.equal(a.prompt_learner.text_prompt.data, b.prompt_learner.text_prompt.data)
    assert torch.equal(a.prompt_learner.img_prompt.data, b.prompt_learner.img_prompt.data)


# ---- two processes, two GPUs, NCCL ------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, out, mode):
    """mode: (shard the text tower?, peer-memory exchanges?)"""
    import torch.distributed as dist
    shard, peer = mode
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    tokens = class_tokens(NCCL_CLASSES)  # 13 classes: parts of 7 and 6
    K, B = NCCL_K, NCCL_B
    model, arch, _ = make_model("small", "fp16", K, tokens, device=f"cuda:{rank}")
    if shard:
        model.shard_text()
    image = synth.make_images(world * B, arch.image_resolution)[rank * B:(rank + 1) * B]
    label = synth.make_labels(world * B, tokens.shape[0])[rank * B:(rank + 1) * B]
    r = StepRunner(model, B, lr=NCCL_LR, process_group=None, world_size=world, peer=peer)
    r.image.copy_(image)
    r.label.copy_(label)
    r.prepare(warmup=2)
    ls = []
    for _ in range(NCCL_STEPS):
        r.step()
        ls.append(r.loss.clone())
    torch.cuda.synchronize()
    torch.save({"loss": torch.stack(ls).cpu(), "tp": model.prompt_learner.text_prompt.data.cpu(),
                "ip": model.prompt_learner.img_prompt.data.cpu(), "collectives": r.collectives},
               os.path.join(out, f"s{int(shard)}p{int(peer is None)}_r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


NCCL_CLASSES = list(range(0, 1000, 77))
NCCL_K, NCCL_B, NCCL_LR, NCCL_STEPS = 8, 4, 0.02, 4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_process_data_parallel_matches_the_oracle(tmp_path):
    """Four SGD steps on 2 GPUs, every combination of {plain data parallelism, class-sharded text tower} x
    {NCCL collectives, peer-memory exchange kernels}, against ONE process running the oracle + torch.optim.SGD on the
    global batch of 2B images (what the two ranks together compute: mean CE over 2B = mean of the ranks' means).
    Replicas stay in lockstep bit-exactly; the variants agree with each other and with the oracle within fp16 rounding."""
    import torch.multiprocessing as mp
    from tests.test_gpu_trajectory import MOM, WD, oracle_trajectory
    world = 2
    modes = [(False, False), (True, False), (False, None), (True, None)]
    for mode in modes:
        mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    res = {(s, p, r): torch.load(tmp_path / f"s{int(s)}p{int(p)}_r{r}.pt") for s, p in
           [(m[0], m[1] is None) for m in modes] for r in range(world)}
    print({k: v["collectives"] for k, v in res.items()})
    for (s, p, r), v in res.items():  # replicas stay in lockstep
        assert torch.equal(v["tp"], res[(s, p, 0)]["tp"]) and torch.equal(v["ip"], res[(s, p, 0)]["ip"])
    # the oracle on the global batch; every step sees the same 2B images (as the workers do)
    tokens = class_tokens(NCCL_CLASSES)
    model, arch, sd = make_model("small", "fp16", NCCL_K, tokens)
    tp0, ip0 = model.prompt_learner.text_prompt.detach().clone(), model.prompt_learner.img_prompt.detach().clone()
    image = synth.make_images(world * NCCL_B, arch.image_resolution)
    label = synth.make_labels(world * NCCL_B, tokens.shape[0])
    batches = [(image, label)] * NCCL_STEPS
    ref_l, ref_tp, ref_ip = oracle_trajectory(sd, "fp16", None, tokens, NCCL_K, tp0, ip0, batches, NCCL_LR)
    tru_l, tru_tp, tru_ip = oracle_trajectory(sd, "fp64", "fp16", tokens, NCCL_K, tp0, ip0, batches, NCCL_LR)
    move_t = (tru_tp - tp0.double().cpu()).abs().max().item()
    move_i = (tru_ip - ip0.double().cpu()).abs().max().item()
    e_ref = ((ref_tp - tru_tp).abs().max().item() / move_t, (ref_ip - tru_ip).abs().max().item() / move_i)
    for (s, p, r), v in res.items():
        if r:
            continue
        # global loss of a step = mean over ranks of the ranks' losses
        gl = torch.stack([res[(s, p, q)]["loss"] for q in range(world)]).mean(0)
        dl = max(abs(float(a) - b) for a, b in zip(gl, ref_l))
        et = (v["tp"].double() - tru_tp).abs().max().item() / move_t
        ei = (v["ip"].double() - tru_ip).abs().max().item() / move_i
        print(f"shard={s} peer={p} [{v['collectives']}]: max |dloss| vs oracle {dl:.3e}; update error / update "
              f"text {et:.3e} img {ei:.3e} (reference's own fp16 SGD: {e_ref[0]:.3e} {e_ref[1]:.3e})")
        assert dl <= 1e-3 * max(1.0, max(abs(x) for x in ref_l))
        ulp_t = float(tp0.abs().max()) * 2 ** -10 / move_t
        ulp_i = float(ip0.abs().max()) * 2 ** -10 / move_i
        assert et <= max(e_ref[0], 0.02) + ulp_t and ei <= max(e_ref[1], 0.02) + ulp_i
    # peer-memory exchanges must have been used where asked for (2 GPUs of one box have P2P)
    assert res[(False, True, 0)]["collectives"] == "peer", res[(False, True, 0)]["collectives"]
