"""Headline benchmark: images/sec of one RPO training step (forward + CE + prompt-gradient backward +
gradient exchange between ranks + SGD update) for CLIP ViT-B/16 with K=24 read-only prompts, 100 classes,
batch 32 per GPU, fp16 (BASELINE.json configs[1]), with the roofline of the masked-attention kernel (and
of every other kernel family of the step), the unmodified-reference-shaped step timed on the SAME GPU under
torch eager, and the CPU reference timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); if started plainly with --gpus N > 1 it
re-launches itself under torch.distributed.run.  Rank 0 prints ONE JSON line.

What the line holds (N = 1): `value` = device-resident throughput of the step (inputs in HBM); `e2e` = the same
through the host API with every batch uploaded from pinned host memory (uint8 pixels, rpo_b200.input_pipeline)
and every step's loss copied back and read by the host while the next step runs (`e2e_sync_loss`: the host blocking on
it in the same step); `trainer_path` = the same through rpo_b200.trainer.RPO.forward_backward (the
drop-in the reference's train.py reaches); `roofline` (masked attention), `roofline_gemm`, `roofline_kernels`;
`gpu_eager_baseline` (the oracle = the reference's op sequence, fp16 torch eager on this GPU; configs 2 and 4);
`cpu_baseline`.  N > 1: `value` = plain data parallelism (the algorithm N = 1 runs: like with like),
`class_sharded` = the class-sharded text tower on the same ranks, `config4` = both again on
the 1000-class shape of BASELINE.json configs[3].
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json `configs`.  configs[0] is the CPU reference's own case (tests/golden).
CONFIGS = {
    2: dict(arch="ViT-B/16", K=24, n_cls=100, batch_per_gpu=32, prec="fp16"),
    3: dict(arch="ViT-L/14", K=24, n_cls=100, batch_per_gpu=16, prec="bf16"),
    4: dict(arch="ViT-B/16", K=24, n_cls=1000, batch_per_gpu=32, prec="fp16"),
    5: dict(arch="ViT-B/16", K=24, n_cls=1000, batch_per_gpu=64, prec="fp16"),
}
K_SWEEP = (4, 8, 16, 24, 48)  # config 5
CONFIG_TEXT = {
    2: "BASELINE.json configs[1]: ViT-B/16, K=24 prompts, 100 synthetic classes, batch 32/GPU, fp16",
    3: "BASELINE.json configs[2]: ViT-L/14, K=24 prompts, 100 synthetic classes, batch 16/GPU, bf16",
    4: "BASELINE.json configs[3]: ViT-B/16, K=24 prompts, 1000 synthetic classes, batch 32/GPU, fp16",
    5: "BASELINE.json configs[4]: ViT-B/16, K sweep {4,8,16,24,48} (value: K=24), 1000 synthetic classes, batch 64/GPU, fp16",
}
WORKLOAD = dict(CONFIGS[2])  # the default; --config replaces it
METRIC = "images_per_sec_train_step_vitb16_k24"
UNIT = "images/s"
STEP_TEXT = "step = fwd + CE + prompt-grad bwd + gradient exchange + SGD(momentum)"
DTYPE_NAME = {"fp16": "fp16", "bf16": "bf16", "fp32": "f32"}


def metric_name(cfg_id):
    return METRIC if cfg_id in (2, 4, 5) else "images_per_sec_train_step_vitl14_k24"


def workload_config(cfg_id, workload):
    """The workload keys of a `config` object."""
    return {"workload": f"{CONFIG_TEXT[cfg_id]}; {STEP_TEXT}", **workload}


L2_POLICY = ("inputs rotate through 8 distinct batches (154 MB > 126 MB L2 at batch 32); a step touches "
             "~1.5 GB of activations")
TEXT_REPLICATED = "replicated on every rank (as the reference)"


def line_config(cfg_id, workload, world, cuda_graph=True, text_tower=TEXT_REPLICATED, collectives=None):
    """The `config` object of the JSON line -- the SAME object in both arms (the reference arm is timed on this arm's
    workload; what its bounded CPU sample departs from is in its `cpu_baseline.sample`)."""
    if collectives is None:
        collectives = "none" if world == 1 else "peer"
    return {**workload_config(cfg_id, workload), "global_batch": workload["batch_per_gpu"] * world,
            "parallelism": f"dp{world}", "l2_policy": L2_POLICY, "cuda_graph": bool(cuda_graph),
            "text_tower": text_tower, "collectives": collectives}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))[key]
        return d["dram_read_bytes"] + d["dram_write_bytes"]
    except Exception:
        return None


def synthetic_tokens(n_cls):
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", "tokens_class1000.npz"))
    t = torch.zeros(n_cls, int(z["context_length"]), dtype=torch.int64)
    t[:, :z["tokens"].shape[1]] = torch.from_numpy(z["tokens"][:n_cls].astype(np.int64))
    return t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 8:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# Reference arms.  The oracle (torch restatement of the unmodified reference, bit-exact vs it on CPU) is
# what can travel to the GPU box (/root/reference does not exist there): "kind" is "port".
# ---------------------------------------------------------------------------------------------------
def oracle_step_fn(workload, batch, device, prec, as_reference=False):
    """One trainer step of the reference (trainers/rpo.py:298-311: loss = model(image, label); zero_grad;
    backward; SGD step; loss.item()) over the oracle.  `as_reference`: additionally keep the attention masks on the
    host, as the reference does (it re-uploads and casts them in every block: clip/model.py:183)."""
    import torch
    from oracle.rpo_oracle import OracleModel, convert_state_dict  # timed baseline (allowed use)
    from rpo_b200 import synth
    arch = synth.ARCHS[workload["arch"]]
    sd = synth.make_state_dict(arch, 0)
    tokens = synthetic_tokens(workload["n_cls"])
    om = OracleModel(convert_state_dict(sd, prec), tokens, workload["K"], prec, device=device)
    if as_reference:
        om.text_mask = om.text_mask.cpu()
        om.visual_mask = om.visual_mask.cpu()
    tp, ip = synth.make_prompt_init(sd, workload["K"])
    image = synth.make_images(batch, arch.image_resolution)
    label = synth.make_labels(batch, workload["n_cls"])
    if torch.device(device).type == "cuda":
        image, label = image.pin_memory(), label.pin_memory()
    tp = tp.to(device, om.dtype).clone().requires_grad_(True)
    ip = ip.to(device, om.dtype).clone().requires_grad_(True)
    opt = torch.optim.SGD([tp, ip], lr=0.01, momentum=0.9, weight_decay=5e-4)

    def step():
        img, lab = image.to(device), label.to(device)   # trainers/rpo.py:318-323
        loss = om.forward(img, tp, ip, lab, training=True)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return float(loss.item())

    return step


def run_cpu_reference(workload, steps, warmup, total_budget_s=200.0, anomaly_too=False):
    """Times the CPU reference on a bounded sample of the workload.  Starts from the full batch and, if
    `steps + warmup` such steps would not fit in `total_budget_s`, shrinks the image batch (the text tower over
    all classes is still paid every step, as in the reference) using the FLOP model t(B) ~ text + 0.079 TF * B."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full_batch = batch = workload["batch_per_gpu"]
    step = oracle_step_fn(workload, batch, "cpu", "fp32")
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    done_warm = 1
    budget = total_budget_s / max(1, steps + warmup)
    text_tf = 0.012 * workload["n_cls"]
    if first > budget:
        full = text_tf + 0.079 * batch
        for b in (16, 8, 4, 2):
            if b >= batch:
                continue
            batch = b
            if first * (text_tf + 0.079 * b) / full <= budget:
                break
        step = oracle_step_fn(workload, batch, "cpu", "fp32")
        done_warm = 0
    for _ in range(max(0, warmup - done_warm)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"{steps} timed steps of fwd+CE+bwd+SGD, {workload['arch']} K={workload['K']} C={workload['n_cls']}, "
              f"batch {batch} of {full_batch}, fp32 (reference PREC=fp32), torch {torch.__version__} CPU, "
              f"{cores} threads, anomaly mode off")
    out = dict(value=batch / dt, unit=UNIT, cores=cores, kind="port", sample=sample)
    if anomaly_too:  # the real trainer switches it on (trainers/rpo.py:288)
        with torch.autograd.set_detect_anomaly(True):
            t0 = time.perf_counter()
            step()
            out["value_anomaly_mode_on"] = batch / (time.perf_counter() - t0)
    return out, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_id = getattr(args, "config", 2)
    cb, dt = run_cpu_reference(WORKLOAD, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": metric_name(cfg_id), "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": line_config(cfg_id, WORKLOAD, max(1, int(args.gpus))),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(workload, steps=10, warmup=3):
    """The reference's own op sequence (oracle = restatement of trainers/rpo.py:161-232 + clip/model.py:181-191,
    bit-identical to the unmodified reference on CPU) in the workload's precision under torch eager on THIS GPU, driven
    like the reference trainer, timed with CUDA events.  Two flavours: tuned (masks resident on the device, anomaly
    mode off) and as the reference ships (host-resident masks re-uploaded per block, anomaly mode on)."""
    import torch
    B = workload["batch_per_gpu"]
    out = {"torch": torch.__version__, "prec": workload["prec"], "steps": steps, "warmup": warmup,
           "what": "oracle (reference-shaped dense forward + autograd backward + torch.optim.SGD + loss.item()) on cuda:0"}

    def timed(step, n):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    step = oracle_step_fn(workload, B, "cuda:0", workload["prec"])
    ms = timed(step, steps)
    out.update(ms_per_step=ms, value=B / ms * 1e3, unit=UNIT, anomaly_mode=False, masks="device-resident")
    del step
    gc.collect()
    torch.cuda.empty_cache()
    step = oracle_step_fn(workload, B, "cuda:0", workload["prec"], as_reference=True)
    with torch.autograd.set_detect_anomaly(True):
        ms2 = timed(step, max(3, steps // 2))
    out["as_shipped"] = {"ms_per_step": ms2, "value": B / ms2 * 1e3, "anomaly_mode": True,
                         "masks": "host-resident, uploaded and cast in every block (clip/model.py:183)"}
    del step
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------
def time_kernel(fn, iters, warm=5):
    """Seconds per launch: `iters` launches captured into one CUDA graph (as the step itself is), so that host-side
    costs (ctypes, TMA descriptor encoding, launch) stay outside the CUDA-event interval."""
    import torch
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3  # seconds per launch


def attention_roofline(arch, K, B, dt, dev, peaks, iters=48):
    """The masked-attention kernel of the metric, timed alone (CUDA events on the launching stream, after warm-up),
    rotating over more distinct buffers than fit in the 126 MB L2.  Algorithmic bytes per launch: SURVEY.md 8(d),
    2 (2L + 2S) hd per (image, head) -- read Q[L], K[S], V[S], write O[L] at 2 B/element."""
    import torch
    from rpo_b200 import _lib
    lib = _lib.load()
    S = (arch.v_res // arch.v_patch) ** 2 + 1
    D, H = arch.v_width, arch.v_heads
    code = _lib.dtype_code(dt)
    per = (B * S * 3 * D + B * K * D + B * (S + K) * D) * 2
    nbuf = max(4, min(12, int(500e6 // per) + 1))
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = [torch.randn(B * S, 3 * D, generator=g, device=dev).to(dt) for _ in range(nbuf)]
    qp = [torch.randn(B * K, D, generator=g, device=dev).to(dt) for _ in range(nbuf)]
    out = [torch.empty(B * (S + K), D, dtype=dt, device=dev) for _ in range(nbuf)]
    if not lib.rpo_ro_attention_fwd_dense_supported(code, S, K, H):
        return None

    def attn(i):
        j = i % nbuf
        _lib.check(lib.rpo_ro_attention_fwd_dense(qkv[j].data_ptr(), qp[j].data_ptr(), out[j].data_ptr(),
                                                  out[j].data_ptr() + B * S * D * 2, B, S, K, H, code,
                                                  _lib.stream_ptr(dev)))

    t = time_kernel(attn, iters)
    L = S + K
    nbytes = 2 * (2 * L + 2 * S) * 64 * H * B
    flops = 4 * L * S * 64 * H * B
    return {
        "kernel": f"ro_attn_fwd_pp / ro_attn_fwd_tc = rpo_ro_attention_fwd_dense (tcgen05; vision tower, one layer: {B} images x "
                  f"{H} heads, L={L} queries, S={S} keys)",
        "bound": "hbm", "achieved": nbytes / t / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
        "frac": nbytes / t / 1e9 / peaks["hbm"], "traffic": ncu_traffic("ro_attn_fwd_tc") if (K, B, S) == (24, 32, 197) else None,
        "peak_source": f"{peaks['source']} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
        "us_per_launch": t * 1e6, "algorithmic_bytes_per_launch": nbytes,
        "tensor_tflops": flops / t / 1e12, "tensor_frac_of_burst": flops / t / 1e12 / peaks["tf_burst"],
        "l2_policy": f"{nbuf} rotated buffer sets of {per / 1e6:.0f} MB (> 126 MB L2 in total)",
    }


def gemm_roofline(arch, K, B, dt, dev, peaks):
    """dominant GEMM: MLP c_fc over all rows of the vision tower, [B(S+K), D] x [4D, D]^T + bias + QuickGELU"""
    import torch
    from rpo_b200 import _lib
    lib = _lib.load()
    S = (arch.v_res // arch.v_patch) ** 2 + 1
    D = arch.v_width
    code = _lib.dtype_code(dt)
    M, N, Kd = B * (S + K), 4 * D, D
    g = torch.Generator(device=dev).manual_seed(0)
    A = [torch.randn(M, Kd, generator=g, device=dev).to(dt) for _ in range(6)]
    Wt = [(torch.randn(N, Kd, generator=g, device=dev) * Kd ** -0.5).to(dt) for _ in range(6)]
    bias = torch.zeros(N, dtype=dt, device=dev)
    Cm = [torch.empty(M, N, dtype=dt, device=dev) for _ in range(6)]  # 6 x (10.9 + 4.7 + 43.5 MB) >> L2

    def gemm(i):
        j = i % 6
        _lib.check(lib.rpo_gemm_bias_act(A[j].data_ptr(), Kd, Wt[j].data_ptr(), Kd, Cm[j].data_ptr(), N, M, N, Kd,
                                         bias.data_ptr(), 1, None, None, None, 0, code, _lib.GEMM_AUTO,
                                         _lib.stream_ptr(dev)))

    t = time_kernel(gemm, 48)
    flops = 2.0 * M * N * Kd
    return {
        "kernel": f"gemm_tc (c_fc + bias + QuickGELU, M={M} N={N} K={Kd})", "bound": "tensor",
        "achieved": flops / t / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
        "frac": flops / t / 1e12 / peaks["tf_burst"], "traffic": ncu_traffic("gemm_fc") if M == 7072 else None,
        "peak_source": f"{peaks['source']} cuBLAS bf16 burst (MEASURED_PEAKS.json bf16_tflops)",
        "us_per_launch": t * 1e6,
    }


def minimal_step_flops(arch, K, B, C, n_c=10.0):
    """FLOPs a step needs when prompts are query-only, the backward covers the prompt
    rows only and the text context is cached (MAC = 2 FLOP)."""
    S = (arch.v_res // arch.v_patch) ** 2 + 1
    L = S + K
    D, E = arch.v_width, arch.embed_dim
    npatch = S - 1
    v_fwd = arch.v_layers * (6 * S * D * D + 2 * K * D * D + 4 * L * S * D + 2 * L * D * D + 16 * L * D * D) \
        + 2 * npatch * D * 3 * arch.v_patch ** 2 + 2 * K * D * E
    v_bwd = arch.v_layers * (20 * K * D * D + 4 * K * S * D) + 2 * K * D * E
    Dt = arch.t_width
    t_min = arch.t_layers * (20 * K * Dt * Dt + 4 * K * n_c * Dt) + 2 * K * Dt * E
    logits = 2 * B * C * K * E * 3
    return B * (v_fwd + v_bwd) + C * 2 * t_min + logits


# ---------------------------------------------------------------------------------------------------
class Job:
    """One workload on this rank: model, StepRunner(s), input pools."""

    def __init__(self, workload, world, rank, dev, pg, shard_text, use_graph=True, peer=None):
        import torch
        from rpo_b200 import synth
        from rpo_b200.clip_weights import SyntheticCLIP
        from rpo_b200.model import CustomCLIP
        self.torch = torch
        self.w, self.world, self.rank, self.dev, self.pg = workload, world, rank, dev, pg
        self.arch = synth.ARCHS[workload["arch"]]
        K, C, B, prec = workload["K"], workload["n_cls"], workload["batch_per_gpu"], workload["prec"]
        sd = synth.make_state_dict(self.arch, 0)
        cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                              INPUT=SimpleNamespace(SIZE=(self.arch.image_resolution,) * 2))
        torch.manual_seed(0)
        self.model = CustomCLIP(cfg, synth.synthetic_classnames(C), "a photo of a _.", SyntheticCLIP(sd, prec),
                                tokens=synthetic_tokens(C), max_batch=B).to(dev)
        self.model.prompt_learner.train()
        self.shard_text = bool(shard_text and world > 1)
        if self.shard_text:
            self.model.shard_text(rank, world, pg)  # each rank runs ceil(C / world) class prompts
        self.use_graph, self.peer = use_graph, peer
        self.B, self.C = B, C
        # inputs: a pool of distinct batches larger than L2 (8 x 19.3 MB at batch 32), different per rank
        self.pool_n = 8
        res = self.arch.image_resolution
        self.pool = [synth.make_images(B, res, seed=1234 + 97 * rank + i) for i in range(self.pool_n)]
        self.labels = [((torch.arange(B) + i + rank) % C).to(torch.int64) for i in range(self.pool_n)]
        self.runner = self._runner(torch.float32)
        self._runner_u8 = None
        self.tp0 = self.model.prompt_learner.text_prompt.data.clone()
        self.ip0 = self.model.prompt_learner.img_prompt.data.clone()

    def _runner(self, image_dtype):
        from rpo_b200.runner import StepRunner
        r = StepRunner(self.model, self.B, lr=0.01, momentum=0.9, weight_decay=5e-4, use_graph=self.use_graph,
                       process_group=self.pg, world_size=self.world, image_dtype=image_dtype, peer=self.peer)
        if image_dtype == self.torch.uint8:
            r.image.copy_(self._u8(0).to(self.dev))
        else:
            r.image.copy_(self.pool[0].to(self.dev))
        r.label.copy_(self.labels[0].to(self.dev))
        return r.prepare(warmup=3)

    def _u8(self, i):
        torch = self.torch
        g = torch.Generator().manual_seed(4321 + 97 * self.rank + i)
        res = self.arch.image_resolution
        return torch.randint(0, 256, (self.B, 3, res, res), generator=g, dtype=torch.uint8)

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def reset(self, runner):
        """every timed leg starts from the same prompts / optimiser state"""
        pl = self.model.prompt_learner
        pl.text_prompt.data.copy_(self.tp0)
        pl.img_prompt.data.copy_(self.ip0)
        runner.mom_buf.zero_()
        runner.first.fill_(1)

    def time_device(self, steps, warmup):
        """device-resident throughput: inputs already in HBM, rotated through the pool"""
        torch, r = self.torch, self.runner
        pool_dev = [p.to(self.dev) for p in self.pool]
        labels_dev = [l.to(self.dev) for l in self.labels]
        self.reset(r)

        def device_step(i):
            r.image.copy_(pool_dev[i % self.pool_n], non_blocking=True)
            r.label.copy_(labels_dev[i % self.pool_n], non_blocking=True)
            r.step()

        for i in range(warmup):
            device_step(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            device_step(i)
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1) * 1e-3, float(r.loss.item())

    def time_e2e(self, steps, warmup, image_dtype, sync_loss=False):
        """end to end through the host API: pinned host -> device upload of every batch (BatchUploader: copy stream,
        two slots) and a device -> host copy of every step's loss, both inside the timed region.  The host reads each
        loss value when the NEXT step is on its way (what rpo_b200.trainer.RPO.forward_backward does by default: the
        value feeds a running average for the log line, trainers/rpo.py:311-313); `sync_loss` = block on it in the
        same step, as a literal `loss.item()` does."""
        torch = self.torch
        from rpo_b200.input_pipeline import BatchUploader
        if image_dtype == torch.float32:
            r = self.runner
        else:
            if self._runner_u8 is None:
                self._runner_u8 = self._runner(image_dtype)
            r = self._runner_u8
        self.reset(r)
        if image_dtype == torch.uint8:
            pool_pin = [self._u8(i).pin_memory() for i in range(self.pool_n)]
        else:
            pool_pin = [p.pin_memory() for p in self.pool]
        labels_pin = [l.pin_memory() for l in self.labels]
        up = BatchUploader(self.dev, self.B, self.arch.image_resolution, image_dtype)
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        landed = [torch.cuda.Event(), torch.cuda.Event()]
        cur = torch.cuda.current_stream()

        def run(n):
            ticket = up.submit(pool_pin[0], labels_pin[0])
            seen = []
            for i in range(n):
                img, lab = up.acquire(ticket)
                r.image.copy_(img, non_blocking=True)
                r.label.copy_(lab, non_blocking=True)
                up.release(ticket)
                r.step()
                # the next batch's upload is enqueued once this step is on its way (copy stream, other slot): it
                # overlaps the step as before, and its host cost no longer sits between the sync and the launch
                nxt = up.submit(pool_pin[(i + 1) % self.pool_n], labels_pin[(i + 1) % self.pool_n]) if i + 1 < n else None
                loss_host[i & 1:(i & 1) + 1].copy_(r.loss.view(1), non_blocking=True)
                landed[i & 1].record(cur)
                j = i if sync_loss else i - 1  # the step whose loss the host reads now
                if j >= 0:
                    landed[j & 1].synchronize()
                    seen.append(float(loss_host[j & 1]))
                ticket = nxt
            if not sync_loss and n > 0:
                landed[(n - 1) & 1].synchronize()
                seen.append(float(loss_host[(n - 1) & 1]))
            assert len(seen) == n  # every step's loss reached the host inside the timed region
            return seen[-1] if seen else 0.0

        run(max(3, warmup))
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1) * 1e-3, up.bytes_per_batch

    def time_trainer(self, steps, warmup):
        """rpo_b200.trainer.RPO.forward_backward (what the reference's train.py reaches through Dassl's run_epoch) with a
        stand-in for the Dassl base class: host batches as a pinned DataLoader hands them over, torch SGD object + lr
        read every step, loss returned every step (lagging one step)."""
        torch = self.torch
        from rpo_b200 import trainer
        t = trainer.RPO.__new__(trainer.RPO)
        t.model, t.device, t.scaler = self.model, self.dev, None
        t.optim = torch.optim.SGD(self.model.prompt_learner.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
        t.batch_idx, t.num_batches = 0, 1 << 30
        t.update_lr = lambda: None
        if not t.fast_path_available():
            return None
        pool_pin = [p.pin_memory() for p in self.pool]
        labels_pin = [l.pin_memory() for l in self.labels]
        pl = self.model.prompt_learner
        pl.text_prompt.data.copy_(self.tp0)
        pl.img_prompt.data.copy_(self.ip0)

        def run(n):
            for i in range(n):
                t.forward_backward({"img": pool_pin[i % self.pool_n], "label": labels_pin[i % self.pool_n]})

        run(max(3, warmup))
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1) * 1e-3

    def close(self):
        self.runner = self._runner_u8 = None
        self.model._engine = None
        self.model = None
        gc.collect()
        self.torch.cuda.empty_cache()


def max_over_ranks(vals, world, dev):
    import torch
    t = torch.tensor([v if v is not None else -1.0 for v in vals], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return [float(x) if x >= 0 else None for x in t]


def measure(workload, world, rank, dev, pg, shard_text, steps, warmup, e2e=True, trainer=False, use_graph=True,
            keep=False):
    """Device-resident (+ end-to-end) throughput of one workload; times are the max over ranks."""
    import torch
    job = Job(workload, world, rank, dev, pg, shard_text, use_graph)
    t_dev, loss = job.time_device(steps, warmup)
    t_e2e = t_e2e_sync = t_e2e32 = t_tr = None
    h2d = 0
    if e2e:
        t_e2e, h2d = job.time_e2e(steps, warmup, torch.uint8)
        t_e2e_sync, _ = job.time_e2e(steps, warmup, torch.uint8, sync_loss=True)
        t_e2e32, h2d32 = job.time_e2e(steps, warmup, torch.float32)
    if trainer:
        t_tr = job.time_trainer(steps, warmup)
    t_dev, t_e2e, t_e2e_sync, t_e2e32, t_tr = max_over_ranks([t_dev, t_e2e, t_e2e_sync, t_e2e32, t_tr], world, dev)
    imgs = workload["batch_per_gpu"] * world * steps
    out = {"value": imgs / t_dev, "ms_per_step": t_dev / steps * 1e3, "images_per_sec_per_gpu": imgs / t_dev / world,
           "loss_after": loss, "collectives": job.runner.collectives,
           "text_tower": (f"class-sharded over {world} ranks (all-gather of text features + reduce-scatter of their "
                          f"gradient)") if job.shard_text else TEXT_REPLICATED,
           "gpu_launches_per_step": job.runner.launches_per_step,
           "device_workspace_bytes": job.runner.eng.device_bytes()}
    if t_e2e is not None:
        out["e2e"] = {"value": imgs / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                      "ms_per_step": t_e2e / steps * 1e3,
                      "note": "pinned uint8 pixels + labels uploaded every step through rpo_b200.input_pipeline.BatchUploader "
                              "(copy stream, two slots), ToTensor + Normalize inside the patch-extraction kernel; every "
                              "step's loss copied to pinned host memory and read by the host while the next step runs "
                              "(the default of rpo_b200.trainer.RPO.forward_backward), the last one before the clock stops"}
        out["e2e_sync_loss"] = {"value": imgs / t_e2e_sync, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                "d2h_bytes_per_step": 4, "ms_per_step": t_e2e_sync / steps * 1e3,
                                "note": "same, the host blocking on every step's loss before it enqueues the next step "
                                        "(a literal loss.item(), trainers/rpo.py:311; RPO_B200_SYNC_LOSS=1 in the trainer)"}
        out["e2e_float32_upload"] = {"value": imgs / t_e2e32, "unit": UNIT, "h2d_bytes_per_step": h2d32,
                                     "d2h_bytes_per_step": 4, "ms_per_step": t_e2e32 / steps * 1e3,
                                     "note": "as e2e, with host-normalised float32 images (what the reference's transform "
                                             "pipeline hands over)"}
    if t_tr is not None:
        out["trainer_path"] = {"value": imgs / t_tr, "unit": UNIT, "ms_per_step": t_tr / steps * 1e3,
                               "note": "rpo_b200.trainer.RPO.forward_backward: BatchUploader + one CUDA-graph replay per "
                                       "step, lr read from the torch optimizer every step, loss returned every step "
                                       "(one step late), pinned float32 host batches"}
    if keep:
        return out, job
    job.close()
    return out


def main_own(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # plain launch with --gpus N: re-exec under torchrun (the driver normally does this itself)
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD if world > 1 else None
    peaks = load_peaks()
    cfg_id = args.config
    W = dict(WORKLOAD)
    K, C, B, prec = W["K"], W["n_cls"], W["batch_per_gpu"], W["prec"]
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[prec]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    main, job = measure(W, world, rank, dev, pg, shard_text=False, steps=args.steps, warmup=args.warmup, e2e=True,
                        trainer=(world == 1), use_graph=not args.no_graph, keep=True)
    clocks = sampler.stop() if sampler else None
    arch = job.model.arch
    job.close()

    line = None
    if rank == 0:
        line = {
            "metric": metric_name(cfg_id), "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAME[prec], "data": "synthetic",
            "config": line_config(cfg_id, W, world, not args.no_graph, main["text_tower"], main["collectives"]),
            "images_per_sec_per_gpu": main["images_per_sec_per_gpu"],
            "clocks": clocks,
            "e2e": main["e2e"], "e2e_sync_loss": main["e2e_sync_loss"],
            "e2e_float32_upload": main["e2e_float32_upload"],
            "gpu_launches": main["gpu_launches_per_step"] * args.steps,
            "gpu_launches_per_step": main["gpu_launches_per_step"],
            "loss_after": main["loss_after"],
            "device_workspace_bytes": main["device_workspace_bytes"],
        }
        if "trainer_path" in main:
            line["trainer_path"] = main["trainer_path"]
        flops = minimal_step_flops(arch, K, B, C)
        step_s = main["ms_per_step"] * 1e-3
        line["step_tensor"] = {"minimal_tflop_per_step_per_gpu": flops / 1e12, "achieved_tflops": flops / step_s / 1e12,
                               "frac_of_sustained_peak": flops / step_s / 1e12 / peaks["tf_sustained"],
                               "peak": peaks["tf_sustained"], "peak_source": peaks["source"]}

    # ---- N > 1: the class-sharded text tower on the same ranks, and both variants on the 1000-class shape ----
    if world > 1 and not args.quick:
        from rpo_b200.text_shard import ClassShard
        extra = {}
        if ClassShard.feasible(C, world):
            extra["class_sharded"] = measure(W, world, rank, dev, pg, True, args.steps, args.warmup, e2e=True)
        if cfg_id == 2:
            W4 = dict(CONFIGS[4])
            c4 = {"config": workload_config(4, W4),
                  "plain_dp": measure(W4, world, rank, dev, pg, False, args.steps, args.warmup, e2e=False),
                  "class_sharded": measure(W4, world, rank, dev, pg, True, args.steps, args.warmup, e2e=False)}
            c4["sharded_over_plain"] = c4["class_sharded"]["value"] / c4["plain_dp"]["value"]
            extra["config4"] = c4
        if rank == 0:
            line.update(extra)

    # ---- N = 1: kernel rooflines, the K sweep of config 5, the reference on this GPU, the CPU baseline ----
    if rank == 0 and world == 1:
        line["roofline"] = attention_roofline(arch, K, B, dt, dev, peaks)
        line["roofline_gemm"] = gemm_roofline(arch, K, B, dt, dev, peaks)
        torch.cuda.empty_cache()
        if not args.quick:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from kernel_bench import bench_kernels
            rows = bench_kernels(prec, W["arch"], B, K, C)
            for r in rows:
                r["share_of_step_note"] = "timed alone (graph of 40 launches, rotated buffers)"
            line["roofline_kernels"] = rows
            torch.cuda.empty_cache()
        if cfg_id == 5:
            sweep = []
            for k in K_SWEEP:
                Wk = dict(W, K=k)
                m = main if k == K else measure(Wk, 1, 0, dev, None, False, args.steps, args.warmup, e2e=False)
                a = attention_roofline(arch, k, B, dt, dev, peaks)
                sweep.append({"K": k, "value": m["value"], "ms_per_step": m["ms_per_step"],
                              "attn_us_per_launch": a["us_per_launch"], "attn_gb_per_s": a["achieved"],
                              "attn_frac_of_hbm_peak": a["frac"], "attn_algorithmic_bytes_per_launch": a["algorithmic_bytes_per_launch"]})
                torch.cuda.empty_cache()
            line["k_sweep"] = sweep
        if not args.quick and not args.no_gpu_baseline:
            eager = {f"config{cfg_id}": gpu_eager_baseline(W)}
            eager[f"config{cfg_id}"]["speedup_of_this_repo"] = main["value"] / eager[f"config{cfg_id}"]["value"]
            if cfg_id == 2:  # BASELINE.md 4.4 names configs 2 and 4 as the bar
                W4 = dict(CONFIGS[4])
                own4 = measure(W4, 1, 0, dev, None, False, args.steps, args.warmup, e2e=False)
                e4 = gpu_eager_baseline(W4, steps=6, warmup=2)
                e4["this_repo"] = {"value": own4["value"], "ms_per_step": own4["ms_per_step"]}
                e4["speedup_of_this_repo"] = own4["value"] / e4["value"]
                eager["config4"] = e4
            line["gpu_eager_baseline"] = eager
        if not args.no_cpu_baseline:
            cb, _ = run_cpu_reference(W, steps=3, warmup=1, total_budget_s=60.0, anomaly_too=True)
            line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (2 = the headline: ViT-B/16 K=24 C=100 B=32 fp16)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="only the step itself (+ the two headline rooflines at N = 1)")
    a = ap.parse_args()
    WORKLOAD.clear()
    WORKLOAD.update(CONFIGS[a.config])
    if a.impl == "reference":
        main_reference(a)
    else:
        main_own(a)
