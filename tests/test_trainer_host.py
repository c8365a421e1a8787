"""Host logic of the trainer plugin mirror (rpo_b200/trainer.py vs reference trainers/rpo.py:235-357)
that needs no GPU: config check, batch parsing, checkpoint loading (format of SURVEY.md H11), and the
N>1 gradient averaging over a world_size-2 gloo group."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn as nn

from rpo_b200 import trainer


class _Prompts(nn.Module):
    def __init__(self, K=3, dt=8, dv=12):
        super().__init__()
        self.text_prompt = nn.Parameter(torch.zeros(K, dt))
        self.img_prompt = nn.Parameter(torch.zeros(K, dv))


def _make_trainer():
    t = trainer.RPO.__new__(trainer.RPO)
    t.device = torch.device("cpu")
    t._models = {"prompt_learner": _Prompts()}
    t.get_model_names = lambda: list(t._models)
    return t


def test_check_cfg_matches_reference():
    t = _make_trainer()
    for prec in ("fp16", "fp32", "amp"):
        t.check_cfg(SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(PREC=prec))))
    with pytest.raises(AssertionError):  # bf16 is an engine extension, not a trainer precision (rpo.py:238)
        t.check_cfg(SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(PREC="bf16"))))


def test_parse_batch_train():
    t = _make_trainer()
    img, lab = t.parse_batch_train({"img": torch.ones(2, 3, 4, 4), "label": torch.tensor([1, 0]), "impath": ["a", "b"]})
    assert img.shape == (2, 3, 4, 4) and lab.dtype == torch.int64


def test_load_model_checkpoint_format(tmp_path):
    """Checkpoint = {"state_dict": {"text_prompt", "img_prompt"}, "epoch"} at
    <dir>/prompt_learner/model.pth.tar-<epoch>; CoOp legacy keys are dropped; loading is non-strict."""
    t = _make_trainer()
    d = tmp_path / "prompt_learner"
    d.mkdir()
    sd = {"text_prompt": torch.full((3, 8), 2.0), "img_prompt": torch.full((3, 12), -1.0),
          "token_prefix": torch.zeros(1), "token_suffix": torch.zeros(1), "unknown_extra": torch.zeros(2)}
    torch.save({"state_dict": sd, "epoch": 15}, d / "model.pth.tar-15")
    t.load_model(str(tmp_path), epoch=15)
    pl = t._models["prompt_learner"]
    assert torch.all(pl.text_prompt == 2.0) and torch.all(pl.img_prompt == -1.0)
    with pytest.raises(FileNotFoundError):
        t.load_model(str(tmp_path), epoch=3)
    t.load_model("", epoch=3)  # skipped, as in the reference


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
    m = _Prompts()
    m.text_prompt.grad = torch.full_like(m.text_prompt, float(rank + 1))
    m.img_prompt.grad = torch.arange(m.img_prompt.numel(), dtype=torch.float32).view_as(m.img_prompt) * (rank + 1)
    w = trainer.allreduce_mean_(list(m.parameters()))
    torch.save({"w": w, "t": m.text_prompt.grad, "i": m.img_prompt.grad}, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_mean_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{k}.pt") for k in range(world)]
    base = torch.arange(36, dtype=torch.float32).view(3, 12)
    for k in range(world):
        assert r[k]["w"] == world
        assert torch.allclose(r[k]["t"], torch.full((3, 8), 1.5))
        assert torch.allclose(r[k]["i"], base * 1.5)


def test_allreduce_mean_is_noop_without_process_group():
    m = _Prompts()
    m.text_prompt.grad = torch.ones_like(m.text_prompt)
    assert trainer.allreduce_mean_(list(m.parameters())) == 1
    assert torch.all(m.text_prompt.grad == 1)


def test_maybe_shard_text_needs_opt_in_and_a_process_group(monkeypatch):
    class M:
        def shard_text(self, group=None):
            raise AssertionError("must not be called")
    monkeypatch.delenv("RPO_B200_SHARD_TEXT", raising=False)
    assert trainer.maybe_shard_text(M()) is None
    monkeypatch.setenv("RPO_B200_SHARD_TEXT", "1")
    assert trainer.maybe_shard_text(M()) is None  # torch.distributed is not initialised here


def _shard_decision_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RPO_B200_SHARD_TEXT"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class M:
        def __init__(self, n_cls):
            self.text_x = torch.zeros(n_cls, 77, 8)
            self.called = False

        def shard_text(self, group=None):
            from rpo_b200.text_shard import ClassShard
            self.called = True
            return ClassShard(self.text_x.shape[0], dist.get_rank(group), dist.get_world_size(group))

    got = {}
    for n_cls in (3, 4, 10, 13):   # over 4 ranks: 3 -> infeasible, 4 -> 1 each, 10 -> 3,3,3,1, 13 -> 4,4,4,1
        m = M(n_cls)
        sh = trainer.maybe_shard_text(m)
        got[n_cls] = None if sh is None else sh.local
    torch.save(got, os.path.join(out, f"d{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_decision_is_the_same_on_every_rank(tmp_path):
    """with fewer classes than needed to give every rank a part, ALL ranks must stay replicated (a
    per-rank decision would leave some ranks inside all_gather / reduce_scatter and others outside: deadlock)."""
    world = 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_shard_decision_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"d{r}.pt") for r in range(world)]
    for n_cls in (3, 4, 10, 13):
        sharded = [g[n_cls] is not None for g in got]
        assert all(sharded) or not any(sharded), (n_cls, sharded)
    assert all(g[3] is None for g in got)
    assert [g[10] for g in got] == [3, 3, 3, 1] and [g[4] for g in got] == [1, 1, 1, 1]
