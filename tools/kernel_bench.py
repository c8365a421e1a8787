"""Per-kernel timings of the step's GEMM shapes and the attention kernels through the C ABI (CUDA events on the
launching stream, buffers rotated so that consecutive launches do not hit L2).  Run on the GPU box:

    python tools/kernel_bench.py [--prec fp16] [--tag note]

Prints one line per kernel: shape, microseconds, TFLOP/s or GB/s, fraction of the measured peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpo_b200 import _lib  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"]
    return 6650.0, 1590.0


USE_GRAPH = True


def timeit(fn, iters=40, warm=5):
    """Seconds per launch.  The `iters` launches are captured into ONE CUDA graph (as the training step is), so
    host-side costs -- ctypes, tensor-map encoding, launch -- are outside the measurement."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if USE_GRAPH:
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", default="fp16")
    ap.add_argument("--tag", default="")
    ap.add_argument("--cublas", action="store_true", help="also time torch.matmul (cuBLAS) on each GEMM shape")
    ap.add_argument("--only", default="", help="comma list: gemm,attn,ln")
    ap.add_argument("--no-dense", action="store_true", help="vision attention through the mma.sync kernel")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches (host-bound for short kernels)")
    ap.add_argument("--nbuf", type=int, default=0, help="override the number of rotated buffers (1 = warm L2)")
    a = ap.parse_args()
    global USE_GRAPH
    USE_GRAPH = not a.no_graph
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.prec]
    lib = _lib.load()
    dev = torch.device("cuda:0")
    code = _lib.dtype_code(dt)
    hbm, tf = peaks()
    g = torch.Generator(device="cpu").manual_seed(0)
    print(f"# kernel_bench {a.tag} prec={a.prec} peaks: {hbm:.0f} GB/s, {tf:.0f} TF/s (measured)")

    # (label, M, N, K, bias, act, residual, gelu_aux)
    shapes = [
        ("v.qkv", 6304, 2304, 768, 1, 0, 0, 0), ("v.q_prompt", 768, 768, 768, 1, 0, 0, 0),
        ("v.out", 7072, 768, 768, 1, 0, 1, 0), ("v.fc", 7072, 3072, 768, 1, 1, 0, 0),
        ("v.proj", 7072, 768, 3072, 1, 0, 1, 0), ("v.patch", 6272, 768, 768, 0, 0, 0, 0),
        ("t.q", 2400, 512, 512, 1, 0, 0, 0), ("t.out", 2400, 512, 512, 1, 0, 1, 0),
        ("t.fc", 2400, 2048, 512, 1, 1, 0, 0), ("t.proj", 2400, 512, 2048, 1, 0, 1, 0),
        ("vb.dpre", 768, 3072, 768, 0, 0, 0, 1), ("vb.dh", 768, 768, 3072, 0, 0, 0, 0),
        ("vb.sq", 768, 768, 768, 0, 0, 0, 0),
        ("tb.dpre", 2400, 2048, 512, 0, 0, 0, 1), ("tb.dh", 2400, 512, 2048, 0, 0, 0, 0),
        ("tb.sq", 2400, 512, 512, 0, 0, 0, 0),
    ]
    only = set(a.only.split(",")) if a.only else {"gemm", "attn", "ln"}
    ws = torch.zeros(lib.rpo_gemm_workspace_bytes(), dtype=torch.uint8, device=dev)  # stream-K workspace (vision stream)
    if "gemm" not in only:
        shapes = []
    for label, M, N, Kd, has_bias, act, has_res, has_aux in shapes:
        per = (M * Kd + N * Kd + M * N * (1 + has_res + has_aux)) * 2
        nbuf = a.nbuf or max(2, min(12, int(300e6 // per) + 1))
        A = [torch.randn(M, Kd, generator=g).to(dt).to(dev) for _ in range(nbuf)]
        W = [(torch.randn(N, Kd, generator=g) * Kd ** -0.5).to(dt).to(dev) for _ in range(nbuf)]
        Cm = [torch.empty(M, N, dtype=dt, device=dev) for _ in range(nbuf)]
        bias = torch.zeros(N, dtype=dt, device=dev) if has_bias else None
        res = [torch.randn(M, N, generator=g).to(dt).to(dev) for _ in range(nbuf)] if has_res else None
        aux = [torch.randn(M, N, generator=g).to(dt).to(dev) for _ in range(nbuf)] if has_aux else None

        def fn(i):
            j = i % nbuf
            _lib.check(lib.rpo_gemm_bias_act_ws(A[j].data_ptr(), Kd, W[j].data_ptr(), Kd, Cm[j].data_ptr(), N, M, N, Kd,
                                                _lib.ptr(bias), act, _lib.ptr(res[j]) if res else None,
                                                _lib.ptr(aux[j]) if aux else None, None, 0, code, _lib.GEMM_AUTO,
                                                ws.data_ptr() if label.startswith("v.") else None,
                                                _lib.stream_ptr(dev)))

        t = timeit(fn)
        fl = 2.0 * M * N * Kd
        # library reference for the same contraction (no epilogue): torch.matmul -> cuBLAS
        tl = timeit(lambda i: torch.matmul(A[i % nbuf], W[i % nbuf].t(), out=Cm[i % nbuf])) if a.cublas else float("nan")
        print(f"gemm {label:10s} M={M:5d} N={N:5d} K={Kd:5d}  {t * 1e6:8.2f} us  {fl / t / 1e12:7.1f} TF/s  "
              f"{fl / t / 1e12 / tf:5.1%} of cuBLAS burst   [cuBLAS same shape: {tl * 1e6:7.2f} us]")
        del A, W, Cm, res, aux

    # attention: vision forward (all rows), vision backward (prompt rows), text prompt-only forward/backward
    for label, G, H, K, n, do_ctx in ([("v.attn_fwd", 32, 12, 24, 197, 1), ("t.attn_fwd", 100, 8, 24, 10, 0)]
                                      if "attn" in only else []):
        D = H * 64
        L = (n if do_ctx else 0) + K
        nbuf = a.nbuf or 10
        qkv = [torch.randn(G * n, 3 * D, generator=g).to(dt).to(dev) for _ in range(nbuf)]
        qp = [torch.randn(G * K, D, generator=g).to(dt).to(dev) for _ in range(nbuf)]
        oc = [torch.empty(G * n, D, dtype=dt, device=dev) for _ in range(nbuf)]
        op = [torch.empty(G * K, D, dtype=dt, device=dev) for _ in range(nbuf)]
        dq = [torch.empty(G * K, D, dtype=dt, device=dev) for _ in range(nbuf)]
        off = torch.arange(0, (G + 1) * n, n, dtype=torch.int32, device=dev)

        dense = do_ctx and not a.no_dense

        def fwd(i):
            j = i % nbuf
            if dense:
                _lib.check(lib.rpo_ro_attention_fwd_dense(qkv[j].data_ptr(), qp[j].data_ptr(), oc[j].data_ptr(),
                                                          op[j].data_ptr(), G, n, K, H, code, _lib.stream_ptr(dev)))
            else:
                _lib.check(lib.rpo_ro_attention_fwd(qkv[j].data_ptr(), qp[j].data_ptr(), oc[j].data_ptr(),
                                                    op[j].data_ptr(), off.data_ptr(), G, K, H, n, 0, do_ctx, code,
                                                    _lib.stream_ptr(dev)))

        def bwd(i):
            j = i % nbuf
            _lib.check(lib.rpo_ro_attention_bwd(qkv[j].data_ptr(), qp[j].data_ptr(), op[j].data_ptr(),
                                                qp[(j + 1) % nbuf].data_ptr(), dq[j].data_ptr(), off.data_ptr(), G, K, H,
                                                n, code, _lib.stream_ptr(dev)))

        t = timeit(fwd)
        by = 2 * (2 * L + 2 * n) * 64 * H * G
        fl = 4.0 * L * n * 64 * H * G
        print(f"attn {label:10s} G={G} H={H} L={L} S={n}  {t * 1e6:8.2f} us  {by / t / 1e9:7.1f} GB/s "
              f"({by / t / 1e9 / hbm:5.1%} of copy peak)  {fl / t / 1e12:6.1f} TF/s")
        t = timeit(bwd)
        by = 2 * (3 * K + 2 * n) * 64 * H * G
        print(f"attn {label.replace('fwd', 'bwd'):10s} G={G} H={H} K={K} S={n}  {t * 1e6:8.2f} us  {by / t / 1e9:7.1f} GB/s "
              f"({by / t / 1e9 / hbm:5.1%} of copy peak)")
        del qkv, qp, oc, op, dq

    # LayerNorm forward
    for rows, D in ([(7072, 768), (2400, 512), (768, 768)] if "ln" in only else []):
        nbuf = a.nbuf or 12
        x = [torch.randn(rows, D, generator=g).to(dt).to(dev) for _ in range(nbuf)]
        y = [torch.empty(rows, D, dtype=dt, device=dev) for _ in range(nbuf)]
        w = torch.ones(D, dtype=torch.float32, device=dev)
        b = torch.zeros(D, dtype=torch.float32, device=dev)

        def ln(i):
            j = i % nbuf
            _lib.check(lib.rpo_layernorm_fwd(x[j].data_ptr(), w.data_ptr(), b.data_ptr(), y[j].data_ptr(), rows, D,
                                             code, _lib.stream_ptr(dev)))

        t = timeit(ln)
        by = 2 * rows * D * 2
        print(f"ln_fwd rows={rows} D={D}  {t * 1e6:8.2f} us  {by / t / 1e9:7.1f} GB/s ({by / t / 1e9 / hbm:5.1%})")


if __name__ == "__main__":
    main()
