"""Phase timeline of CTA 0 of the tcgen05 attention kernel (diagnostics build: RPO_DIAG=1 python -m rpo_b200.build).

    RPO_DIAG=1 python -m rpo_b200.build --force && python tools/attn_trace.py [--arch ViT-B/16] [--batch 32] [--K 24]

Events per work item (SM clock, relative to the first event): producer K/V issued, Q issued; MMA thread: S issued,
P-full seen, PV issued; softmax warp 4: S-full seen, scores in registers, maxima exchanged, P stored; epilogue warp 12:
O-full seen, stores done."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpo_b200 import _lib  # noqa: E402

EV = {10: "q_issue", 0: "s_issue", 3: "sm_sfull", 4: "sm_ld1", 9: "sm_folded", 5: "sm_maxx", 11: "sm_blk2", 12: "sm_blk4",
      13: "sm_blk6", 14: "sm_blocks", 15: "sm_stwait", 6: "sm_pdone", 1: "pv_pfull", 2: "pv_issued", 7: "ep_ofull", 8: "ep_done"}
ORDER = [10, 0, 3, 4, 9, 5, 11, 12, 13, 14, 15, 6, 1, 2, 7, 8]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="ViT-B/16")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--K", type=int, default=24)
    ap.add_argument("--prec", default="fp16")
    a = ap.parse_args()
    lib = _lib.load()
    if not hasattr(lib, "rpo_diag_set_attn_trace"):
        raise SystemExit("not a diagnostics build: RPO_DIAG=1 python -m rpo_b200.build --force")
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.prec]
    dev = torch.device("cuda:0")
    H, n = (12, 197) if a.arch == "ViT-B/16" else (16, 257)
    D, G, K = H * 64, a.batch, a.K
    qkv = torch.randn(G * n, 3 * D, device=dev).to(dt)
    qp = torch.randn(G * K, D, device=dev).to(dt)
    oc = torch.empty(G * n, D, dtype=dt, device=dev)
    op = torch.empty(G * K, D, dtype=dt, device=dev)
    trace = torch.zeros(64 * 16 + 4 * 148 + 64, dtype=torch.int64, device=dev)
    code = _lib.dtype_code(dt)

    def run():
        _lib.check(lib.rpo_ro_attention_fwd_dense(qkv.data_ptr(), qp.data_ptr(), oc.data_ptr(), op.data_ptr(), G, n, K, H,
                                                  code, _lib.stream_ptr(dev)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"20 back-to-back eager launches: {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per launch")
    lib.rpo_diag_set_attn_trace.argtypes = [C.c_void_p]
    _lib.check(lib.rpo_diag_set_attn_trace(trace.data_ptr()))
    # a CUDA graph of back-to-back launches (as the step is): the stamps are those of the last launch, so the
    # dependency timing is the steady state and the host's launch cost is out of the picture
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(6):
            run()
    g.replay()
    torch.cuda.synchronize()
    trace.zero_()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph of 6 launches: {e0.elapsed_time(e1) / 6 * 1e3:.2f} us per launch")
    _lib.check(lib.rpo_diag_set_attn_trace(None))
    wall = trace.cpu()[1024:1024 + 4 * 148].view(148, 4)
    t = trace.cpu()[:1024].view(64, 16)
    nz = t[t > 0]
    t0 = int(nz.min())
    print(f"kernel entry {int(t[0, 12]) - t0}, prologue done {int(t[0, 13]) - t0}, all roles done {int(t[0, 14]) - t0}, "
          f"tmem released {int(t[0, 15]) - t0} (SM clocks)")
    w = wall[wall[:, 0] > 0].double()
    if len(w):  # the two-slot kernel stamps every CTA with the global timer
        w0 = float(w[:, 0].min())
        print(f"{len(w)} CTAs, wall clock (us after the first entry): entry {float(w[:, 0].min()) - w0:.2f} .. "
              f"{(float(w[:, 0].max()) - w0) / 1e3:.2f}; dependency released {(float(w[:, 1].min()) - w0) / 1e3:.2f} .. "
              f"{(float(w[:, 1].max()) - w0) / 1e3:.2f}; roles done {(float(w[:, 2].min()) - w0) / 1e3:.2f} .. "
              f"{(float(w[:, 2].max()) - w0) / 1e3:.2f}; exit {(float(w[:, 3].min()) - w0) / 1e3:.2f} .. "
              f"{(float(w[:, 3].max()) - w0) / 1e3:.2f}")
    order = [e for e in ORDER if int(t[1:, e].max()) > 0]  # the two kernels record different sub-phases
    print("tile  " + " ".join(f"{EV[e]:>10s}" for e in order))
    for j in range(64):
        if int(t[j].max()) == 0:
            break
        print(f"{j:4d}  " + " ".join(f"{(int(t[j, e]) - t0) if t[j, e] > 0 else -1:10d}" for e in order))


if __name__ == "__main__":
    main()
