"""Phase timeline of the single-CTA tcgen05 GEMM on a small-M shape (tuning aid): a CUDA graph of back-to-back launches
(as in the step), the last one traced through RPO_GEMM_TRACE; prints mean cycles per phase over the CTAs."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpo_b200 import _lib
M, N, Kd = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (768, 768, 768))]
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(0)
nb = 6
A = [torch.randn(M, Kd, generator=g).half().to(dev) for _ in range(nb)]
W = [(torch.randn(N, Kd, generator=g) * Kd ** -0.5).half().to(dev) for _ in range(nb)]
Cm = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(nb)]
trace = torch.zeros(4096, 16, dtype=torch.int64, device=dev)
RES = [torch.randn(M, N, generator=g).half().to(dev) for _ in range(nb)] if os.environ.get("TRACE_RES") else None
BIAS = torch.randn(N, generator=g).half().to(dev)
ACT = int(os.environ.get("TRACE_ACT", "0"))


def call(j):
    _lib.check(lib.rpo_gemm_bias_act(A[j].data_ptr(), Kd, W[j].data_ptr(), Kd, Cm[j].data_ptr(), N, M, N, Kd, BIAS.data_ptr(),
                                     ACT, RES[j].data_ptr() if RES else None, None, None, 0, 1, _lib.GEMM_AUTO,
                                     _lib.stream_ptr(dev)))


for j in range(nb):
    call(j)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for j in range(nb - 1):
        call(j)
    os.environ["RPO_GEMM_TRACE"] = hex(trace.data_ptr())
    call(nb - 1)
    del os.environ["RPO_GEMM_TRACE"]
gr.replay()
torch.cuda.synchronize()
trace.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
gr.replay()
e1.record()
torch.cuda.synchronize()
t = trace.cpu().double()
t = t[t[:, 1] > 0]
print(f"M={M} N={N} K={Kd}: {len(t)} CTAs traced; graph of {nb} launches: {e0.elapsed_time(e1) * 1e3 / nb:.2f} us per launch")
names = ["setup (barriers, TMEM alloc, sync)", "dependency wait (PDL)", "first operands landed", "main loop -> accumulator complete",
         "epilogue of the first tile", "rest + teardown sync"]
for i, nm in enumerate(names):
    print(f"   {nm:40s} {float((t[:, i + 2] - t[:, i + 1]).mean()):8.0f} cycles")
print(f"   CTA lifetime {float((t[:, 7] - t[:, 1]).mean()):.0f} cycles;  kernel span (first entry -> last exit, globaltimer) "
      f"{float(t[:, 8].max() - t[:, 0].min()) / 1e3:.2f} us;  entry skew {float(t[:, 0].max() - t[:, 0].min()) / 1e3:.2f} us")
