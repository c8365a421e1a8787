"""Per-kernel timings of the step's GEMM shapes, the attention kernels and LayerNorm through the C ABI (CUDA events
on the launching stream, launches captured into one CUDA graph as the step is, buffers rotated so that consecutive
launches do not hit L2).  Run on the GPU box:

    python tools/kernel_bench.py [--prec fp16] [--cublas] [--arch ViT-B/16] [--batch 32] [--K 24] [--classes 100]

Prints one line per kernel: shape, microseconds, TFLOP/s or GB/s, fraction of the measured peak.  bench.py imports
`bench_kernels` for its `roofline_kernels` list.  Algorithmic bytes / FLOPs per launch follow SURVEY.md 8(d)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpo_b200 import _lib  # noqa: E402

ARCH_DIMS = {  # Dv, Hv, S (context rows), patch K extent (padded), Dt, Ht
    "ViT-B/16": dict(Dv=768, S=197, pk=768, Dt=512),
    "ViT-L/14": dict(Dv=1024, S=257, pk=640, Dt=768),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def timeit(fn, iters=40, warm=5, use_graph=True):
    """Seconds per launch.  The `iters` launches are captured into ONE CUDA graph (as the training step is), so
    host-side costs -- ctypes, tensor-map encoding, launch -- are outside the measurement."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def gemm_shapes(arch, B, K, C, n_c=10):
    d = ARCH_DIMS[arch]
    Dv, S, pk, Dt = d["Dv"], d["S"], d["pk"], d["Dt"]
    Mv, Mp, Mt = B * (S + K), B * K, C * K
    # (label, M, N, K, bias, act, residual, gelu_aux)
    return [
        ("v.qkv", B * S, 3 * Dv, Dv, 1, 0, 0, 0), ("v.out", Mv, Dv, Dv, 1, 0, 1, 0), ("v.fc", Mv, 4 * Dv, Dv, 1, 1, 0, 0),
        ("v.proj", Mv, Dv, 4 * Dv, 1, 0, 1, 0), ("v.patch", B * (S - 1), Dv, pk, 0, 0, 0, 0),
        ("t.q", Mt, Dt, Dt, 1, 0, 0, 0), ("t.out", Mt, Dt, Dt, 1, 0, 1, 0), ("t.fc", Mt, 4 * Dt, Dt, 1, 1, 0, 0),
        ("t.proj", Mt, Dt, 4 * Dt, 1, 0, 1, 0),
        ("vb.dpre", Mp, 4 * Dv, Dv, 0, 0, 0, 1), ("vb.dh", Mp, Dv, 4 * Dv, 0, 0, 0, 0), ("vb.sq", Mp, Dv, Dv, 0, 0, 0, 0),
        ("tb.dpre", Mt, 4 * Dt, Dt, 0, 0, 0, 1), ("tb.dh", Mt, Dt, 4 * Dt, 0, 0, 0, 0), ("tb.sq", Mt, Dt, Dt, 0, 0, 0, 0),
    ]


def bench_kernels(prec="fp16", arch="ViT-B/16", B=32, K=24, C=100, only=("gemm", "attn", "ln"), cublas=False,
                  nbuf_override=0, dense=True, use_graph=True, gemm_labels=None, device="cuda:0"):
    """Returns a list of dicts: kernel, shape, us, bound, achieved, peak, unit, frac (+ cublas_us for GEMMs)."""
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[prec]
    lib = _lib.load()
    dev = torch.device(device)
    code = _lib.dtype_code(dt)
    hbm, tf, _ = peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    d = ARCH_DIMS[arch]
    rows = []

    def randn(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g, device=dev) * scale).to(dt)

    for label, M, N, Kd, has_bias, act, has_res, has_aux in (gemm_shapes(arch, B, K, C) if "gemm" in only else []):
        if gemm_labels and label not in gemm_labels:
            continue
        per = (M * Kd + N * Kd + M * N * (1 + has_res + has_aux)) * 2
        nbuf = nbuf_override or max(2, min(12, int(300e6 // per) + 1))
        A = [randn(M, Kd) for _ in range(nbuf)]
        W = [randn(N, Kd, scale=Kd ** -0.5) for _ in range(nbuf)]
        Cm = [torch.empty(M, N, dtype=dt, device=dev) for _ in range(nbuf)]
        bias = torch.zeros(N, dtype=dt, device=dev) if has_bias else None
        res = [randn(M, N) for _ in range(nbuf)] if has_res else None
        aux = [randn(M, N) for _ in range(nbuf)] if has_aux else None

        def fn(i):
            j = i % nbuf
            _lib.check(lib.rpo_gemm_bias_act(A[j].data_ptr(), Kd, W[j].data_ptr(), Kd, Cm[j].data_ptr(), N, M, N, Kd,
                                             _lib.ptr(bias), act, _lib.ptr(res[j]) if res else None,
                                             _lib.ptr(aux[j]) if aux else None, None, 0, code, _lib.GEMM_AUTO,
                                             _lib.stream_ptr(dev)))

        t = timeit(fn, use_graph=use_graph)
        fl = 2.0 * M * N * Kd
        # library reference for the same contraction (no epilogue): torch.matmul -> cuBLAS
        tl = timeit(lambda i: torch.matmul(A[i % nbuf], W[i % nbuf].t(), out=Cm[i % nbuf]), use_graph=use_graph) \
            if cublas else None
        rows.append(dict(kernel=f"gemm {label}", shape=f"M={M} N={N} K={Kd}", us=t * 1e6, bound="tensor",
                         achieved=fl / t / 1e12, peak=tf, unit="TFLOP/s", frac=fl / t / 1e12 / tf, flops=fl,
                         cublas_us=tl * 1e6 if tl else None))
        del A, W, Cm, res, aux

    # attention: vision forward (all rows), vision backward (prompt rows), text prompt-only forward/backward
    attn_cases = [("v.attn", B, d["Dv"] // 64, K, d["S"], 1), ("t.attn", C, d["Dt"] // 64, K, 10, 0)]
    for label, G, H, Kp, n, do_ctx in (attn_cases if "attn" in only else []):
        D = H * 64
        L = (n if do_ctx else 0) + Kp
        per = (G * n * 3 * D + G * Kp * D * 3 + G * n * D) * 2
        nbuf = nbuf_override or max(2, min(10, int(450e6 // per) + 1))
        qkv = [randn(G * n, 3 * D) for _ in range(nbuf)]
        qp = [randn(G * Kp, D) for _ in range(nbuf)]
        oc = [torch.empty(G * n, D, dtype=dt, device=dev) for _ in range(nbuf)]
        op = [torch.empty(G * Kp, D, dtype=dt, device=dev) for _ in range(nbuf)]
        dq = [torch.empty(G * Kp, D, dtype=dt, device=dev) for _ in range(nbuf)]
        off = torch.arange(0, (G + 1) * n, n, dtype=torch.int32, device=dev)
        use_dense = bool(do_ctx and dense)

        def fwd(i):
            j = i % nbuf
            if use_dense:
                _lib.check(lib.rpo_ro_attention_fwd_dense(qkv[j].data_ptr(), qp[j].data_ptr(), oc[j].data_ptr(),
                                                          op[j].data_ptr(), G, n, Kp, H, code, _lib.stream_ptr(dev)))
            else:
                _lib.check(lib.rpo_ro_attention_fwd(qkv[j].data_ptr(), qp[j].data_ptr(), oc[j].data_ptr(),
                                                    op[j].data_ptr(), off.data_ptr(), G, Kp, H, n, 0, do_ctx, code,
                                                    _lib.stream_ptr(dev)))

        def bwd(i):
            j = i % nbuf
            _lib.check(lib.rpo_ro_attention_bwd(qkv[j].data_ptr(), qp[j].data_ptr(), op[j].data_ptr(),
                                                qp[(j + 1) % nbuf].data_ptr(), dq[j].data_ptr(), off.data_ptr(), G, Kp,
                                                H, n, code, _lib.stream_ptr(dev)))

        t = timeit(fwd, use_graph=use_graph)
        by = 2 * (2 * L + 2 * n) * 64 * H * G
        fl = 4.0 * L * n * 64 * H * G
        rows.append(dict(kernel=f"attn {label}_fwd" + ("" if not do_ctx else (" (tcgen05)" if use_dense else " (mma.sync)")),
                         shape=f"G={G} H={H} L={L} S={n}", us=t * 1e6, bound="hbm", achieved=by / t / 1e9, peak=hbm,
                         unit="GB/s", frac=by / t / 1e9 / hbm, bytes=by, flops=fl, tensor_tflops=fl / t / 1e12))
        t = timeit(bwd, use_graph=use_graph)
        by = 2 * (3 * Kp + 2 * n) * 64 * H * G
        rows.append(dict(kernel=f"attn {label}_bwd", shape=f"G={G} H={H} K={Kp} S={n}", us=t * 1e6, bound="hbm",
                         achieved=by / t / 1e9, peak=hbm, unit="GB/s", frac=by / t / 1e9 / hbm, bytes=by))
        del qkv, qp, oc, op, dq

    # LayerNorm forward / backward
    ln_cases = [(B * (d["S"] + K), d["Dv"]), (C * K, d["Dt"]), (B * K, d["Dv"])]
    for nrows, D in (ln_cases if "ln" in only else []):
        nbuf = nbuf_override or 12
        x = [randn(nrows, D) for _ in range(nbuf)]
        y = [torch.empty(nrows, D, dtype=dt, device=dev) for _ in range(nbuf)]
        w = torch.ones(D, dtype=torch.float32, device=dev)
        b = torch.zeros(D, dtype=torch.float32, device=dev)

        def ln(i):
            j = i % nbuf
            _lib.check(lib.rpo_layernorm_fwd(x[j].data_ptr(), w.data_ptr(), b.data_ptr(), y[j].data_ptr(), nrows, D,
                                             code, _lib.stream_ptr(dev)))

        def lnb(i):
            j = i % nbuf
            _lib.check(lib.rpo_layernorm_bwd(x[j].data_ptr(), x[(j + 1) % nbuf].data_ptr(), w.data_ptr(),
                                             x[(j + 2) % nbuf].data_ptr(), y[j].data_ptr(), nrows, D, code,
                                             _lib.stream_ptr(dev)))

        t = timeit(ln, use_graph=use_graph)
        by = 2 * nrows * D * 2
        rows.append(dict(kernel="ln_fwd", shape=f"rows={nrows} D={D}", us=t * 1e6, bound="hbm", achieved=by / t / 1e9,
                         peak=hbm, unit="GB/s", frac=by / t / 1e9 / hbm, bytes=by))
        if nrows != ln_cases[0][0]:  # the backward only ever runs over prompt rows
            t = timeit(lnb, use_graph=use_graph)
            by = 4 * nrows * D * 2  # dy, x, residual gradient in; dx out
            rows.append(dict(kernel="ln_bwd", shape=f"rows={nrows} D={D}", us=t * 1e6, bound="hbm",
                             achieved=by / t / 1e9, peak=hbm, unit="GB/s", frac=by / t / 1e9 / hbm, bytes=by))
    return rows


def format_row(r):
    s = f"{r['kernel']:28s} {r['shape']:28s} {r['us']:8.2f} us  {r['achieved']:8.1f} {r['unit']:8s} {r['frac']:6.1%} of peak"
    if r.get("cublas_us"):
        s += f"   [cuBLAS same shape: {r['cublas_us']:7.2f} us -> {r['cublas_us'] / r['us']:.2f}x]"
    if r.get("tensor_tflops"):
        s += f"   {r['tensor_tflops']:6.1f} TF/s"
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", default="fp16")
    ap.add_argument("--arch", default="ViT-B/16")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--K", type=int, default=24)
    ap.add_argument("--classes", type=int, default=100)
    ap.add_argument("--tag", default="")
    ap.add_argument("--cublas", action="store_true", help="also time torch.matmul (cuBLAS) on each GEMM shape")
    ap.add_argument("--only", default="", help="comma list: gemm,attn,ln")
    ap.add_argument("--gemms", default="", help="comma list of GEMM labels (v.qkv, t.q, ...)")
    ap.add_argument("--no-dense", action="store_true", help="vision attention through the mma.sync kernel")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches (host-bound for short kernels)")
    ap.add_argument("--nbuf", type=int, default=0, help="override the number of rotated buffers (1 = warm L2)")
    a = ap.parse_args()
    hbm, tf, src = peaks()
    print(f"# kernel_bench {a.tag} prec={a.prec} arch={a.arch} B={a.batch} K={a.K} C={a.classes} "
          f"peaks: {hbm:.0f} GB/s, {tf:.0f} TF/s ({src})")
    only = tuple(a.only.split(",")) if a.only else ("gemm", "attn", "ln")
    for r in bench_kernels(a.prec, a.arch, a.batch, a.K, a.classes, only, a.cublas, a.nbuf, not a.no_dense,
                           not a.no_graph, set(a.gemms.split(",")) if a.gemms else None):
        print(format_row(r), flush=True)


if __name__ == "__main__":
    main()
