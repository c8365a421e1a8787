"""Phase timeline of the tcgen05 attention CTAs (tuning aid): launches the kernel once with RPO_ATTN_TRACE set to a
device buffer and prints the mean SM-clock cycles each CTA spends per phase."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpo_b200 import _lib
G, H, K, n = 32, 12, 24, 197
D = H * 64
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(0)
qkv = torch.randn(G * n, 3 * D, generator=g).half().to(dev)
qp = torch.randn(G * K, D, generator=g).half().to(dev)
oc = torch.empty(G * n, D, dtype=torch.float16, device=dev)
op = torch.empty(G * K, D, dtype=torch.float16, device=dev)
ncta = 2 * 148
trace = torch.zeros(ncta, 8, dtype=torch.int64, device=dev)
for warm in range(3):
    _lib.check(lib.rpo_ro_attention_fwd_dense(qkv.data_ptr(), qp.data_ptr(), oc.data_ptr(), op.data_ptr(), G, n, K, H, 1,
                                              _lib.stream_ptr(dev)))
torch.cuda.synchronize()
os.environ["RPO_ATTN_TRACE"] = hex(trace.data_ptr())
_lib.check(lib.rpo_ro_attention_fwd_dense(qkv.data_ptr(), qp.data_ptr(), oc.data_ptr(), op.data_ptr(), G, n, K, H, 1,
                                          _lib.stream_ptr(dev)))
torch.cuda.synchronize()
t = trace.cpu().double()
names = ["setup (barriers, TMEM alloc, sync)", "loads Q+K landed", "S MMA done (seen by softmax)", "pass 1 (max) done",
         "pass 2 + P stored + all warps arrived", "PV MMA done (seen by softmax)", "epilogue + teardown sync"]
order = [0, 1, 2, 3, 4, 5, 6, 7]
for tile in (0, 1):
    tt = t[tile::2]  # first item of CTA b is item b: even CTAs start on tile 0, odd on tile 1
    print(f"tile {tile}: total {float((tt[:, 7] - tt[:, 0]).mean()):.0f} cycles")
    for i, nm in enumerate(names):
        a, b = order[i], order[i + 1]
        print(f"   {nm:42s} {float((tt[:, b] - tt[:, a]).mean()):8.0f}")
