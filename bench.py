"""Headline benchmark: images/sec of one RPO training step (forward + CE + prompt-gradient backward +
gradient all-reduce + SGD update) for CLIP ViT-B/16 with K=24 read-only prompts, 100 classes,
batch 32 per GPU, fp16 (BASELINE.json configs[1]), plus the roofline of the masked-attention and
dominant GEMM kernels and the CPU reference timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); if started plainly with --gpus N > 1 it
re-launches itself under torch.distributed.run.  Rank 0 prints ONE JSON line.
"""
import argparse
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

WORKLOAD = dict(arch="ViT-B/16", K=24, n_cls=100, batch_per_gpu=32, prec="fp16")
METRIC = "images_per_sec_train_step_vitb16_k24"
UNIT = "images/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, refreshed by tools/profile_step.sh runs); None if absent."""
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
# CPU reference arm: the oracle port (torch restatement of the unmodified reference, bit-exact vs it
# on CPU) on the host cores.  /root/reference does not exist on the GPU box, so "kind" is "port".
# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(batch):
    import torch
    from oracle.rpo_oracle import OracleModel, convert_state_dict  # timed CPU baseline (allowed use)
    from rpo_b200 import synth
    arch = synth.ARCHS[WORKLOAD["arch"]]
    sd = synth.make_state_dict(arch, 0)
    tokens = synthetic_tokens(WORKLOAD["n_cls"])
    # PREC=fp32 is the reference's own CPU-friendly precision (trainers/rpo.py:247-249); fp16 on CPU is
    # emulated and several times slower, which would flatter the GPU arm
    om = OracleModel(convert_state_dict(sd, "fp32"), tokens, WORKLOAD["K"], "fp32", device="cpu")
    tp, ip = synth.make_prompt_init(sd, WORKLOAD["K"])
    image = synth.make_images(batch, arch.image_resolution)
    label = synth.make_labels(batch, WORKLOAD["n_cls"])
    tp = tp.clone().requires_grad_(True)
    ip = ip.clone().requires_grad_(True)
    opt = torch.optim.SGD([tp, ip], lr=0.01, momentum=0.9, weight_decay=5e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = om.forward(image, tp, ip, label, training=True)
        loss.backward()
        opt.step()
        return float(loss.item())

    return step


def run_cpu_reference(steps, warmup, total_budget_s=200.0):
    """Times the CPU reference on a bounded sample of the workload.  Starts from the full batch (32)
    and, if `steps + warmup` such steps would not fit in `total_budget_s`, shrinks the image batch
    (the text tower over all 100 classes is still paid every step, as in the reference) using the
    FLOP model t(B) ~ 1.2 TF (text) + 0.079 TF * B."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = WORKLOAD["batch_per_gpu"]
    step = cpu_reference_step_fn(batch)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    done_warm = 1
    budget = total_budget_s / max(1, steps + warmup)
    if first > budget:
        full = 1.2 + 0.079 * batch
        for b in (16, 8, 4, 2):
            batch = b
            if first * (1.2 + 0.079 * b) / full <= budget:
                break
        step = cpu_reference_step_fn(batch)
        done_warm = 0
    for _ in range(max(0, warmup - done_warm)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f"{steps} timed steps of fwd+CE+bwd+SGD, ViT-B/16 K=24 C=100, batch {batch} of 32, fp32 "
              f"(reference PREC=fp32), torch {torch.__version__} CPU, {cores} threads")
    return dict(value=batch / dt, unit=UNIT, cores=cores, kind="port", sample=sample), dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = run_cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload="ViT-B/16 K=24 C=100 train step (CPU reference, oracle port)", **WORKLOAD),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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


def kernel_rooflines(model, peaks):
    """Times the two kernels the metric names in isolation (CUDA events on the launching stream, after
    warm-up), rotating over more distinct buffers than fit in the 126 MB L2, and relates them to the
    measured peaks.  Algorithmic bytes/FLOPs per launch follow SURVEY.md 8(d)."""
    import torch
    from rpo_b200 import _lib
    lib = _lib.load()
    dev = model.w_mm.device
    arch, K, B = model.arch, model.K, WORKLOAD["batch_per_gpu"]
    S = (arch.v_res // arch.v_patch) ** 2 + 1
    D, H = arch.v_width, arch.v_heads
    dt = model.dtype
    code = _lib.dtype_code(dt)
    nbuf = 12  # 12 x (29 MB qkv + 1.2 MB qp + 10.9 MB out) >> L2
    g = torch.Generator(device="cpu").manual_seed(0)
    qkv = [(torch.randn(B * S, 3 * D, generator=g) * 1.0).to(dt).to(dev) for _ in range(nbuf)]
    qp = [(torch.randn(B * K, D, generator=g)).to(dt).to(dev) for _ in range(nbuf)]
    out = [torch.empty(B * (S + K), D, dtype=dt, device=dev) for _ in range(nbuf)]
    off = torch.arange(0, (B + 1) * S, S, dtype=torch.int32, device=dev)

    def attn(i):
        j = i % nbuf
        _lib.check(lib.rpo_ro_attention_fwd_dense(qkv[j].data_ptr(), qp[j].data_ptr(), out[j].data_ptr(),
                                                  out[j].data_ptr() + B * S * D * 2, B, S, K, H, code,
                                                  _lib.stream_ptr(dev)))

    t_attn = time_kernel(attn, 48)
    L = S + K
    attn_bytes = 2 * (2 * L + 2 * S) * 64 * H * B  # read Q[L], K[S], V[S]; write O[L]; 2 B/elem
    attn_flops = 4 * L * S * 64 * H * B
    roof_attn = {
        "kernel": "ro_attn_fwd_tc = rpo_ro_attention_fwd_dense (tcgen05; vision, per layer: 32 images x 12 heads, L=221 queries, S=197 keys)",
        "bound": "hbm", "achieved": attn_bytes / t_attn / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
        "frac": attn_bytes / t_attn / 1e9 / peaks["hbm"], "traffic": ncu_traffic("ro_attn_fwd_tc"),
        "peak_source": f"{peaks['source']} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
        "us_per_launch": t_attn * 1e6, "algorithmic_bytes_per_launch": attn_bytes,
        "tensor_tflops": attn_flops / t_attn / 1e12, "tensor_frac_of_burst": attn_flops / t_attn / 1e12 / peaks["tf_burst"],
    }
    del qkv, qp, out
    # dominant GEMM: MLP c_fc over all rows of the vision tower, [7072,768] x [3072,768]^T + bias + QuickGELU
    M, N, Kd = B * (S + K), 4 * D, D
    A = [(torch.randn(M, Kd, generator=g)).to(dt).to(dev) for _ in range(6)]
    Wt = [(torch.randn(N, Kd, generator=g) * Kd ** -0.5).to(dt).to(dev) for _ in range(6)]
    bias = torch.zeros(N, dtype=dt, device=dev)
    Cm = [torch.empty(M, N, dtype=dt, device=dev) for _ in range(6)]  # 6 x (10.9 + 4.7 + 43.5 MB) >> L2

    def gemm(i):
        j = i % 6
        _lib.check(lib.rpo_gemm_bias_act(A[j].data_ptr(), Kd, Wt[j].data_ptr(), Kd, Cm[j].data_ptr(), N, M, N, Kd,
                                         bias.data_ptr(), 1, None, None, None, 0, code, _lib.GEMM_AUTO,
                                         _lib.stream_ptr(dev)))

    t_gemm = time_kernel(gemm, 48)
    flops = 2.0 * M * N * Kd
    roof_gemm = {
        "kernel": "gemm_tc (c_fc + bias + QuickGELU, M=7072 N=3072 K=768)", "bound": "tensor",
        "achieved": flops / t_gemm / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
        "frac": flops / t_gemm / 1e12 / peaks["tf_burst"], "traffic": ncu_traffic("gemm_fc"),
        "peak_source": f"{peaks['source']} cuBLAS bf16 burst (MEASURED_PEAKS.json bf16_tflops)",
        "us_per_launch": t_gemm * 1e6,
    }
    return roof_attn, roof_gemm


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

    from rpo_b200 import _lib, synth
    from rpo_b200.clip_weights import SyntheticCLIP
    from rpo_b200.model import CustomCLIP
    from rpo_b200.runner import StepRunner

    peaks = load_peaks()
    arch = synth.ARCHS[WORKLOAD["arch"]]
    K, C, B, prec = WORKLOAD["K"], WORKLOAD["n_cls"], WORKLOAD["batch_per_gpu"], WORKLOAD["prec"]
    sd = synth.make_state_dict(arch, 0)
    cfg = SimpleNamespace(TRAINER=SimpleNamespace(RPO=SimpleNamespace(K=K, PREC=prec)),
                          INPUT=SimpleNamespace(SIZE=(arch.image_resolution,) * 2))
    torch.manual_seed(0)
    model = CustomCLIP(cfg, synth.synthetic_classnames(C), "a photo of a _.", SyntheticCLIP(sd, prec),
                       tokens=synthetic_tokens(C), max_batch=B).to(dev)
    model.prompt_learner.train()
    pg = dist.group.WORLD if world > 1 else None
    shard_text = world > 1 and not args.no_shard_text
    pipeline = args.pipeline
    if shard_text:
        model.shard_text(rank, world, pg)  # each rank runs ceil(C / world) class prompts
    runner = StepRunner(model, B, lr=0.01, momentum=0.9, weight_decay=5e-4, use_graph=not args.no_graph,
                        process_group=pg, world_size=world, pipeline=pipeline)

    # inputs: a pool of distinct batches larger than L2 (8 x 19.3 MB), different per rank
    pool_n = 8
    pool = [synth.make_images(B, arch.image_resolution, seed=1234 + 97 * rank + i) for i in range(pool_n)]
    labels = [((torch.arange(B) + i + rank) % C).to(torch.int64) for i in range(pool_n)]
    pool_dev = [p.to(dev) for p in pool]
    labels_dev = [l.to(dev) for l in labels]
    pool_pin = [p.pin_memory() for p in pool]
    labels_pin = [l.pin_memory() for l in labels]
    runner.image.copy_(pool_dev[0])
    runner.label.copy_(labels_dev[0])
    runner.prepare(warmup=3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        runner.image.copy_(pool_dev[i % pool_n], non_blocking=True)
        runner.label.copy_(labels_dev[i % pool_n], non_blocking=True)
        runner.step()

    # ---- device-resident throughput (`value`) ----
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(i)
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    loss_after = float(runner.loss.item())

    # ---- end to end through the host API: pinned host -> device copy of every batch inside the timed
    # region, loss read back to the host every step ----
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    stage = [torch.empty_like(runner.image) for _ in range(2)]
    stage_lab = [torch.empty_like(runner.label) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            stage[s].copy_(pool_pin[i % pool_n], non_blocking=True)
            stage_lab[s].copy_(labels_pin[i % pool_n], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_run(n):
        cur = torch.cuda.current_stream()
        for s in range(2):
            consumed[s].record(cur)
        upload(0)
        last = 0.0
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                upload(i + 1)  # overlaps with step i
            cur.wait_event(ready[s])
            runner.image.copy_(stage[s], non_blocking=True)
            runner.label.copy_(stage_lab[s], non_blocking=True)
            consumed[s].record(cur)
            runner.step()
            loss_host.copy_(runner.loss.view(1), non_blocking=True)
            cur.synchronize()  # the reference reads loss.item() every step (trainers/rpo.py:311)
            last = float(loss_host[0])
        return last

    e2e_run(max(3, args.warmup))
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop() if sampler else None

    times = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(times[0]), float(times[1])

    if rank == 0:
        imgs = B * world * args.steps
        h2d = B * 3 * arch.image_resolution ** 2 * 4 + B * 8
        line = {
            "metric": METRIC, "value": imgs / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {
                "workload": "BASELINE.json configs[1]: ViT-B/16, K=24 prompts, 100 synthetic classes, batch 32/GPU, "
                            "fp16; step = fwd + CE + prompt-grad bwd + allreduce + SGD(momentum)",
                **WORKLOAD, "global_batch": B * world, "parallelism": f"dp{world}",
                "l2_policy": "inputs rotate through 8 distinct batches (154 MB > 126 MB L2); a step touches ~1.5 GB "
                             "of activations",
                "cuda_graph": not args.no_graph,
                "pipeline": ("context rows (cls + patches: independent of the prompts) of batch i+1 run on a second "
                             "stream beside the prompt-row chain / backward / SGD of batch i; one batch in flight, every "
                             "timed step completes one batch; same trajectory as the sequential step") if pipeline
                else "none: each step runs one batch start to end",
                "text_tower": (f"class-sharded over {world} ranks (all-gather of text features + reduce-scatter of "
                               f"their gradient)") if shard_text else "replicated on every rank (as the reference)",
            },
            "images_per_sec_per_gpu": imgs / t_dev / world,
            "clocks": clocks,
            "e2e": {"value": imgs / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": t_e2e / args.steps * 1e3,
                    "note": "pinned fp32 images+labels uploaded every step (double-buffered on a copy stream), loss "
                            "read back and synchronised every step"},
            "gpu_launches": runner.launches_per_step * args.steps,
            "gpu_launches_per_step": runner.launches_per_step,
            "loss_after": loss_after,
            "device_workspace_bytes": runner.eng.device_bytes(),
        }
        flops = minimal_step_flops(model.arch, K, B, C)
        step_s = t_dev / args.steps
        line["step_tensor"] = {"minimal_tflop_per_step_per_gpu": flops / 1e12, "achieved_tflops": flops / step_s / 1e12,
                               "frac_of_sustained_peak": flops / step_s / 1e12 / peaks["tf_sustained"],
                               "peak": peaks["tf_sustained"], "peak_source": peaks["source"]}
        roof_attn, roof_gemm = kernel_rooflines(model, peaks)
        line["roofline"] = roof_attn
        line["roofline_gemm"] = roof_gemm
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = run_cpu_reference(steps=1, warmup=1)
            line["cpu_baseline"] = cb
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
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shard-text", action="store_true",
                    help="N > 1: run every class prompt on every rank (as the reference) instead of class-sharding "
                         "the text tower over the ranks")
    ap.add_argument("--pipeline", action="store_true",
                    help="overlap the context rows of the next batch with the prompt-row chain of the current one "
                         "(StepRunner(pipeline=True); measured slower on B200, see DESIGN.md)")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_own(a)
