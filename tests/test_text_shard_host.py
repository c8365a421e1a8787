"""Host logic of the class-sharded text tower (rpo_b200/text_shard.py) on CPU: the class
partition, and -- over a world_size-2 gloo group -- that all-gather of text features + reduce-scatter
of their gradient + all-reduce / world of the flat prompt gradient reproduces the gradient of the
global-batch mean loss computed in one process.  The towers are stand-in differentiable functions
(the real ones only exist as CUDA kernels); the logits / CE block is the reference's
(trainers/rpo.py:215-230)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn.functional as F

from rpo_b200.text_shard import ClassShard, TextExchange

C_, K_, E_, DT, DV, B_ = 5, 3, 8, 6, 7, 4  # 5 classes over 2 ranks: parts of 3 and 2, padded to 6


def test_class_shard_partition():
    for n_cls, world in ((100, 8), (1000, 8), (5, 2), (7, 7), (8, 3), (1, 1)):
        shards = [ClassShard(n_cls, r, world) for r in range(world)]
        covered = [c for s in shards for c in range(s.first, s.first + s.local)]
        assert covered == list(range(n_cls))
        assert all(s.per == shards[0].per and s.n_pad == s.per * world >= n_cls for s in shards)
        assert all(1 <= s.local <= s.per for s in shards)
    # 10 classes over 8 ranks (e.g. EuroSAT on 8 GPUs): parts of 2 would leave ranks 5..7 empty.  The refusal must
    # come from EVERY rank -- a rank that stays replicated while others issue the collectives deadlocks them.
    for r in range(8):
        with pytest.raises(ValueError):
            ClassShard(10, r, 8)
    assert not ClassShard.feasible(10, 8) and not ClassShard.feasible(5, 4) and ClassShard.feasible(8, 8)
    assert ClassShard.feasible(100, 8) and ClassShard.feasible(13, 2) and not ClassShard.feasible(9, 8)
    with pytest.raises(ValueError):
        ClassShard(10, 2, 2)


def _towers(seed=0):
    g = torch.Generator().manual_seed(seed)
    Wt = torch.randn(C_, DT, E_, generator=g)   # per-class stand-in "text tower"
    Wv = torch.randn(DV, E_, generator=g)
    img = torch.randn(2 * B_, DV, generator=g)  # global batch of 2*B_ "images"
    lab = torch.arange(2 * B_) % C_
    tp = torch.randn(K_, DT, generator=g)
    ip = torch.randn(K_, DV, generator=g)
    return Wt, Wv, img, lab, tp, ip


def _text_feat(tp, Wt):  # [C, K, E]
    return torch.tanh(torch.einsum("kd,cde->cke", tp, Wt))


def _img_feat(ip, img, Wv):  # [B, K, E]
    return torch.tanh((img[:, None, :] + ip[None]) @ Wv)


def _loss(text_f, img_f, lab):
    text_f = text_f / text_f.norm(dim=-1, keepdim=True)
    img_f = img_f / img_f.norm(dim=-1, keepdim=True)
    logits = 14.0 * torch.einsum("bke,cke->bc", img_f, text_f) / K_
    return F.cross_entropy(logits, lab)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Wt, Wv, img, lab, tp, ip = _towers()
    shard = ClassShard(C_, rank, world)
    ex = TextExchange(shard, K_, E_, torch.float32, "cpu")
    tp = tp.clone().requires_grad_(True)
    ip = ip.clone().requires_grad_(True)
    # forward: local classes, all-gather
    tf_local = _text_feat(tp, Wt[shard.slice])
    ex.text_feat[ex.r0:ex.r0 + ex.nl] = tf_local.detach().reshape(-1, E_)
    ex.gather_text_features()
    tf_all = ex.text_feat[:C_ * K_].clone().view(C_, K_, E_).requires_grad_(True)
    mine = slice(rank * B_, (rank + 1) * B_)
    loss = _loss(tf_all, _img_feat(ip, img[mine], Wv), lab[mine])  # mean over the LOCAL batch
    # backward: logits block, reduce-scatter, local text tower, image side
    loss.backward()
    ex.d_text_feat[:C_ * K_] = tf_all.grad.reshape(-1, E_)
    ex.scatter_text_grads()
    tf_local.backward(ex.d_text_feat[ex.r0:ex.r0 + ex.nl].view(shard.local, K_, E_))
    flat = torch.cat([tp.grad.reshape(-1), ip.grad.reshape(-1)])
    dist.all_reduce(flat)
    flat /= world
    assert torch.all(ex.text_feat[C_ * K_:] == 0) and torch.all(ex.d_text_feat[C_ * K_:] == 0)  # padding rows
    torch.save({"flat": flat, "loss": loss.detach(), "text_feat": ex.text_feat.clone()}, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_matches_single_process_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{k}.pt") for k in range(world)]
    Wt, Wv, img, lab, tp, ip = _towers()
    tp = tp.clone().requires_grad_(True)
    ip = ip.clone().requires_grad_(True)
    tf = _text_feat(tp, Wt)
    loss = _loss(tf, _img_feat(ip, img, Wv), lab)  # mean over the global batch
    loss.backward()
    want = torch.cat([tp.grad.reshape(-1), ip.grad.reshape(-1)])
    for k in range(world):
        assert torch.allclose(r[k]["text_feat"][:C_ * K_], tf.detach().reshape(-1, E_), atol=1e-6)
        assert torch.allclose(r[k]["flat"], want, rtol=1e-5, atol=1e-6), (r[k]["flat"] - want).abs().max()
    assert torch.allclose((r[0]["loss"] + r[1]["loss"]) / 2, loss.detach(), atol=1e-6)


def test_exchange_is_a_noop_at_world_1():
    ex = TextExchange(ClassShard(4, 0, 1), 2, 3, torch.float32, "cpu")
    ex.text_feat.fill_(1.0)
    ex.d_text_feat.fill_(2.0)
    ex.gather_text_features()
    ex.scatter_text_grads()
    assert torch.all(ex.text_feat == 1.0) and torch.all(ex.d_text_feat == 2.0)


def test_class_shard_partition_properties():
    """every admissible (n_cls, world): contiguous, disjoint, covering, equal padded parts; otherwise a loud error"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(1, 2000), st.integers(1, 16))
    def check(n_cls, world):
        per = -(-n_cls // world)
        empty_ranks = [r for r in range(world) if r * per >= n_cls]
        assert ClassShard.feasible(n_cls, world) == (not empty_ranks)
        if empty_ranks:  # the same (loud) answer on every rank, empty or not
            for r in range(world):
                with pytest.raises(ValueError):
                    ClassShard(n_cls, r, world)
            return
        nxt = 0
        for r in range(world):
            s = ClassShard(n_cls, r, world)
            assert s.first == nxt and s.per == per and s.n_pad == per * world and 1 <= s.local <= per
            assert s.slice == slice(s.first, s.first + s.local)
            nxt = s.first + s.local
        assert nxt == n_cls

    check()
