"""CPU: the numbers bench.py reports against are the ones SURVEY.md 8(d) states (algorithmic FLOPs of a step,
algorithmic bytes of the masked-attention kernel), the class partition bench.py uses at N > 1 covers config 2's
classes, and the reference arm prints the contract's JSON keys (run here on a tiny stand-in workload)."""
import json

import pytest

import bench
from rpo_b200 import synth
from rpo_b200.text_shard import ClassShard


def _arch(a):
    """the fields bench.minimal_step_flops reads (engine naming), from a synth.ClipArch"""
    from types import SimpleNamespace
    return SimpleNamespace(v_res=a.image_resolution, v_patch=a.vision_patch_size, v_width=a.vision_width,
                           v_layers=a.vision_layers, embed_dim=a.embed_dim, t_width=a.transformer_width,
                           t_layers=a.transformer_layers)


def test_minimal_step_flops_matches_survey():
    arch = _arch(synth.ARCHS["ViT-B/16"])
    # 38.72 + 3.59 GFLOP per image (vision forward + prompt-row backward), 1.528 GFLOP per class
    # forward and the same again backward; 1.660 TFLOP per config-2 step in total
    assert bench.minimal_step_flops(arch, 24, 1, 0) / 1e9 == pytest.approx(38.72 + 3.59, rel=2e-3)
    assert bench.minimal_step_flops(arch, 24, 0, 1) / 2 / 1e9 == pytest.approx(1.528, rel=2e-3)
    assert bench.minimal_step_flops(arch, 24, 32, 100) / 1e12 == pytest.approx(1.660, rel=2e-3)


def test_attention_algorithmic_bytes_match_survey():
    # bytes = 2 (2L + 2S) hd per (image, head, layer); K=24 -> 15.41 MB per image over 12 heads x 12
    # layers; one launch = 32 images x 12 heads of one layer = 41.09 MB (bench.py::kernel_rooflines)
    S, K, H, B = 197, 24, 12, 32
    L = S + K
    per_launch = 2 * (2 * L + 2 * S) * 64 * H * B
    assert per_launch == 41091072
    assert per_launch * 12 / B / 1e6 == pytest.approx(15.41, rel=1e-3)
    assert 4 * L * S * 64 * H * 12 / 1e9 == pytest.approx(1.605, rel=1e-3)  # GFLOP per image


def test_class_partition_of_the_bench_workload():
    C = bench.WORKLOAD["n_cls"]
    for world in (2, 4, 8):
        parts = [ClassShard(C, r, world) for r in range(world)]
        assert sum(p.local for p in parts) == C and parts[0].per * world >= C


def test_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    monkeypatch.setitem(bench.WORKLOAD, "arch", "tiny")
    monkeypatch.setitem(bench.WORKLOAD, "K", 2)
    monkeypatch.setitem(bench.WORKLOAD, "n_cls", 3)
    monkeypatch.setitem(bench.WORKLOAD, "batch_per_gpu", 2)
    monkeypatch.delenv("RANK", raising=False)
    from types import SimpleNamespace
    bench.main_reference(SimpleNamespace(steps=1, warmup=1, gpus=1))
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the `config` object is the own arm's, key for key (bench.line_config builds it for both)
    assert line["config"] == bench.line_config(2, bench.WORKLOAD, 1)
    assert line["config"]["global_batch"] == 2 and line["config"]["parallelism"] == "dp1"
    # ranks other than 0 exit without work
    monkeypatch.setenv("RANK", "1")
    bench.main_reference(SimpleNamespace(steps=1, warmup=1, gpus=2))
    assert capsys.readouterr().out == ""
