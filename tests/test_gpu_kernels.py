"""GPU parity of every hot-path kernel, called through the C ABI (librpo_b200.so via ctypes), against
plain PyTorch fp32 references of the same op (floating-point kernels -> torch reference, tolerance
stated per test: 1e-5 relative for f32, 1e-3 relative-to-max for f16, 8e-3 for bf16's 8-bit mantissa).
"""
import ctypes as C

import pytest
import torch

from rpo_b200 import _lib

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
TOL = {"fp32": 1e-5, "fp16": 1e-3, "bf16": 8e-3}


def dev():
    return torch.device("cuda:0")


def st():
    return _lib.stream_ptr(dev())


def relmax(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def randn(*shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev())


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("rows,D", [(1, 128), (37, 192), (1000, 512), (7072, 768), (333, 1024)])
def test_layernorm_fwd_bwd(prec, rows, D):
    lib = _lib.load()
    dt = DT[prec]
    x = randn(rows, D, dtype=dt, seed=1, scale=2.0) + 0.5
    w = (1 + 0.1 * torch.randn(D, generator=torch.Generator().manual_seed(2))).to(dev())
    b = (0.1 * torch.randn(D, generator=torch.Generator().manual_seed(3))).to(dev())
    y = torch.empty_like(x)
    _lib.check(lib.rpo_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), rows, D,
                                     _lib.dtype_code(dt), st()))
    xr = x.float().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), w, b, 1e-5)
    assert relmax(y, yr.detach().to(dt)) <= TOL[prec]
    dy = randn(rows, D, dtype=dt, seed=4)
    dres = randn(rows, D, dtype=dt, seed=5)
    dx = torch.empty_like(x)
    _lib.check(lib.rpo_layernorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), dres.data_ptr(), dx.data_ptr(), rows,
                                     D, _lib.dtype_code(dt), st()))
    yr.backward(dy.float())
    ref = xr.grad + dres.float()
    assert relmax(dx, ref) <= TOL[prec] * 2
    _lib.check(lib.rpo_layernorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), None, dx.data_ptr(), rows, D,
                                     _lib.dtype_code(dt), st()))
    assert relmax(dx, xr.grad) <= TOL[prec] * 2


# ---------------------------------------------------------------------------------------------------
def quickgelu(x):
    return x * torch.sigmoid(1.702 * x)


def quickgelu_grad(z):
    s = torch.sigmoid(1.702 * z)
    return s * (1 + 1.702 * z * (1 - s))


def run_gemm(A, B, prec, backend, bias=None, act=0, residual=None, gelu_aux=None, aux_row0=None):
    lib = _lib.load()
    M, Kd = A.shape
    N = B.shape[0]
    Cm = torch.zeros(M, N, dtype=A.dtype, device=A.device)
    aux = torch.zeros(M - aux_row0, N, dtype=A.dtype, device=A.device) if aux_row0 is not None else None
    _lib.check(lib.rpo_gemm_bias_act(
        A.data_ptr(), Kd, B.data_ptr(), Kd, Cm.data_ptr(), N, M, N, Kd, _lib.ptr(bias), act, _lib.ptr(residual),
        _lib.ptr(gelu_aux), _lib.ptr(aux), aux_row0 or 0, _lib.dtype_code(A.dtype), backend, st()))
    torch.cuda.synchronize()
    return Cm, aux


def ref_gemm(A, B, bias=None, act=0, residual=None, gelu_aux=None):
    v = A.float() @ B.float().t()
    if bias is not None:
        v = v + bias.float()
    pre = v.clone()
    if act:
        v = quickgelu(v)
    if gelu_aux is not None:
        v = v * quickgelu_grad(gelu_aux.float())
    if residual is not None:
        v = v + residual.float()
    return v, pre


GEMM_SHAPES = [(7072, 768, 768), (300, 2304, 768), (768, 3072, 768), (768, 768, 3072), (1, 512, 512),
               (129, 1536, 512), (2400, 128, 2048), (6272, 768, 768), (100, 96, 64),
               # CTA-pair (cta_group::2) tiles: 256 x 256 and 256 x 128, ragged M tails in the second CTA
               (6304, 2304, 768), (10000, 384, 256), (9999, 128, 64), (2400, 2048, 512)]


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("M,N,Kd", GEMM_SHAPES)
def test_gemm_tcgen05_plain(prec, M, N, Kd):
    dt = DT[prec]
    A = randn(M, Kd, dtype=dt, seed=10)
    B = randn(N, Kd, dtype=dt, seed=11, scale=Kd ** -0.5)
    out, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05)
    ref, _ = ref_gemm(A, B)
    assert relmax(out, ref) <= TOL[prec]
    # the two backends see identical inputs and accumulate in f32: they agree to rounding
    out2, _ = run_gemm(A, B, prec, _lib.GEMM_SIMT)
    assert relmax(out, out2) <= TOL[prec]


@pytest.mark.parametrize("prec,backend", [("fp32", _lib.GEMM_SIMT), ("fp16", _lib.GEMM_SIMT),
                                          ("fp16", _lib.GEMM_TCGEN05), ("bf16", _lib.GEMM_TCGEN05)])
def test_gemm_epilogues(prec, backend):
    dt = DT[prec]
    M, N, Kd = 333, 256, 192
    A = randn(M, Kd, dtype=dt, seed=20)
    B = randn(N, Kd, dtype=dt, seed=21, scale=Kd ** -0.5)
    bias = randn(N, dtype=dt, seed=22, scale=0.3)
    res = randn(M, N, dtype=dt, seed=23)
    auxg = randn(M, N, dtype=dt, seed=24)
    tol = TOL[prec] * 3
    # bias + QuickGELU, with the pre-activation of rows >= 100 captured
    out, aux = run_gemm(A, B, prec, backend, bias=bias, act=1, aux_row0=100)
    ref, pre = ref_gemm(A, B, bias=bias, act=1)
    assert relmax(out, ref) <= tol
    assert relmax(aux, pre[100:]) <= tol
    # bias + residual
    out, _ = run_gemm(A, B, prec, backend, bias=bias, residual=res)
    ref, _ = ref_gemm(A, B, bias=bias, residual=res)
    assert relmax(out, ref) <= tol
    # gelu-gradient epilogue of the backward
    out, _ = run_gemm(A, B, prec, backend, gelu_aux=auxg)
    ref, _ = ref_gemm(A, B, gelu_aux=auxg)
    assert relmax(out, ref) <= tol


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_gemm_pair_epilogues(prec):
    """The CTA-pair kernel at the vision tower's real shapes, every fused epilogue."""
    dt = DT[prec]
    tol = TOL[prec] * 3
    M = 7072
    # c_fc: bias + QuickGELU, pre-activation of the prompt rows (>= 6304) captured
    A = randn(M, 768, dtype=dt, seed=50)
    B = randn(3072, 768, dtype=dt, seed=51, scale=768 ** -0.5)
    bias = randn(3072, dtype=dt, seed=52, scale=0.3)
    out, aux = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, bias=bias, act=1, aux_row0=6304)
    ref, pre = ref_gemm(A, B, bias=bias, act=1)
    assert relmax(out, ref) <= tol
    assert relmax(aux, pre[6304:]) <= tol
    # c_proj: bias + residual, long K
    A = randn(M, 3072, dtype=dt, seed=53)
    B = randn(768, 3072, dtype=dt, seed=54, scale=3072 ** -0.5)
    bias = randn(768, dtype=dt, seed=55, scale=0.3)
    res = randn(M, 768, dtype=dt, seed=56)
    out, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, bias=bias, residual=res)
    ref, _ = ref_gemm(A, B, bias=bias, residual=res)
    assert relmax(out, ref) <= tol
    # 256 x 128 pair tiles with the gelu-gradient epilogue
    A = randn(10000, 256, dtype=dt, seed=57)
    B = randn(384, 256, dtype=dt, seed=58, scale=256 ** -0.5)
    auxg = randn(10000, 384, dtype=dt, seed=59)
    out, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, gelu_aux=auxg)
    ref, _ = ref_gemm(A, B, gelu_aux=auxg)
    assert relmax(out, ref) <= tol


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_gemm_dynamic_schedule_bit_identical(prec, monkeypatch):
    """Dynamic tile schedule of the single-CTA kernel (one CTA per tile, later tiles taken over through cluster launch
    control) against the static round-robin schedule: which CTA computes a tile must not change a single bit.  Shapes
    with more tiles than SMs: 128 x 128 / 128 x 64 / 128 x 32 tiles; the pair-kernel shapes (always static) ride along
    as a repeatability check."""
    dt = DT[prec]
    for (M, N, Kd, with_res, act) in [(7072, 3072, 768, False, 1), (7072, 768, 3072, True, 0), (6304, 2304, 768, False, 0),
                                      (7072, 768, 768, True, 0), (20000, 384, 256, False, 0), (30000, 192, 128, False, 0),
                                      (40000, 96, 64, True, 0)]:
        A = randn(M, Kd, dtype=dt, seed=80)
        B = randn(N, Kd, dtype=dt, seed=81, scale=Kd ** -0.5)
        bias = randn(N, dtype=dt, seed=82, scale=0.3)
        res = randn(M, N, dtype=dt, seed=83) if with_res else None
        monkeypatch.setenv("RPO_GEMM_DYNAMIC", "0")
        static, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, bias=bias, act=act, residual=res)
        monkeypatch.setenv("RPO_GEMM_DYNAMIC", "2")
        for rep in range(3):
            dyn, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, bias=bias, act=act, residual=res)
            assert torch.equal(static, dyn), (M, N, Kd, rep)
        ref, _ = ref_gemm(A, B, bias=bias, act=act, residual=res)
        assert relmax(dyn, ref) <= TOL[prec] * 3, (M, N, Kd)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_gemm_cluster_splitk_epilogues(prec):
    """Long-K, small-M problems take the cluster split-K kernel (partials reduced through distributed shared
    memory): every epilogue, ragged M, K ranges that do not divide evenly."""
    dt = DT[prec]
    tol = TOL[prec] * 3
    for (M, N, Kd) in [(300, 128, 2048), (768, 768, 3072), (2400, 512, 2048), (77, 64, 2112)]:
        A = randn(M, Kd, dtype=dt, seed=80)
        B = randn(N, Kd, dtype=dt, seed=81, scale=Kd ** -0.5)
        bias = randn(N, dtype=dt, seed=82, scale=0.3)
        res = randn(M, N, dtype=dt, seed=83)
        auxg = randn(M, N, dtype=dt, seed=84)
        out, aux = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, bias=bias, act=1, aux_row0=M // 3)
        ref, pre = ref_gemm(A, B, bias=bias, act=1)
        assert relmax(out, ref) <= tol, (M, N, Kd)
        assert relmax(aux, pre[M // 3:]) <= tol
        out, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, bias=bias, residual=res)
        ref, _ = ref_gemm(A, B, bias=bias, residual=res)
        assert relmax(out, ref) <= tol, (M, N, Kd)
        out, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, gelu_aux=auxg)
        ref, _ = ref_gemm(A, B, gelu_aux=auxg)
        assert relmax(out, ref) <= tol, (M, N, Kd)
        out2, _ = run_gemm(A, B, prec, _lib.GEMM_TCGEN05, gelu_aux=auxg)
        assert torch.equal(out, out2)  # fixed reduction order


def test_gemm_f32_exact():
    """RPO_F32 uses true fp32 FMAs (no tf32): 1e-5 relative to the fp64 result."""
    M, N, Kd = 257, 130, 777
    A = randn(M, Kd, dtype=torch.float32, seed=30)
    B = randn(N, Kd, dtype=torch.float32, seed=31)
    out, _ = run_gemm(A, B, "fp32", _lib.GEMM_AUTO)
    ref = (A.double() @ B.double().t()).float()
    assert relmax(out, ref) <= 1e-5


def test_gemm_rejects_tcgen05_for_f32():
    A = randn(128, 64, dtype=torch.float32, seed=1)
    with pytest.raises(_lib.RpoError):
        run_gemm(A, A, "fp32", _lib.GEMM_TCGEN05)


# ---------------------------------------------------------------------------------------------------
def ref_attention(qkv, qp, off, K, H, causal, do_ctx):
    """fp32 reference with an explicit additive mask shaped like the reference's
    (trainers/rpo.py:140-159): prompt columns are -inf for every row."""
    D = H * 64
    G = len(off) - 1
    out_ctx = torch.zeros(qkv.shape[0], D)
    out_p = torch.zeros(G * K, D)
    for g in range(G):
        n = off[g + 1] - off[g]
        blk = qkv[off[g]:off[g + 1]].float()
        q = torch.cat([blk[:, :D], qp[g * K:(g + 1) * K].float()])  # [n+K, D]
        k = torch.cat([blk[:, D:2 * D], torch.zeros(K, D)])
        v = torch.cat([blk[:, 2 * D:], torch.zeros(K, D)])
        L = n + K
        mask = torch.zeros(L, L)
        if causal:
            mask = torch.full((L, L), float("-inf")).triu_(1)
        mask[:, n:] = float("-inf")
        qh = q.view(L, H, 64).transpose(0, 1)
        kh = k.view(L, H, 64).transpose(0, 1)
        vh = v.view(L, H, 64).transpose(0, 1)
        p = torch.softmax(qh @ kh.transpose(1, 2) / 8.0 + mask, dim=-1)
        o = (p @ vh).transpose(0, 1).reshape(L, D)
        out_ctx[off[g]:off[g + 1]] = o[:n]
        out_p[g * K:(g + 1) * K] = o[n:]
    return out_ctx, out_p


ATT_CASES = [
    # (name, context lengths, K, H, causal)
    ("vision", [197] * 3, 24, 12, 0),
    ("vision_k4", [197] * 2, 4, 12, 0),
    ("vitl", [257] * 2, 24, 16, 0),
    ("text", [9, 10, 11, 30, 1, 53], 24, 8, 1),
    ("tiny", [17, 17], 5, 2, 0),
]


@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("case", ATT_CASES, ids=[c[0] for c in ATT_CASES])
def test_ro_attention_fwd_bwd(prec, case):
    lib = _lib.load()
    _, lens, K, H, causal = case
    dt = DT[prec]
    D = H * 64
    G = len(lens)
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    Mc = off[-1]
    qkv = randn(Mc, 3 * D, dtype=dt, seed=40)
    qp = randn(G * K, D, dtype=dt, seed=41)
    off_d = torch.tensor(off, dtype=torch.int32, device=dev())
    out_ctx = torch.zeros(Mc, D, dtype=dt, device=dev())
    out_p = torch.zeros(G * K, D, dtype=dt, device=dev())
    code = _lib.dtype_code(dt)
    _lib.check(lib.rpo_ro_attention_fwd(qkv.data_ptr(), qp.data_ptr(), out_ctx.data_ptr(), out_p.data_ptr(),
                                        off_d.data_ptr(), G, K, H, max(lens), causal, 1, code, st()))
    rc, rp = ref_attention(qkv.cpu(), qp.cpu(), off, K, H, causal, True)
    tol = TOL[prec] * 2
    assert relmax(out_ctx.cpu(), rc) <= tol
    assert relmax(out_p.cpu(), rp) <= tol
    # prompt-only pass (text tower per step): same prompt rows, context output untouched
    out_p2 = torch.zeros_like(out_p)
    sentinel = torch.full_like(out_ctx, 7.0)
    _lib.check(lib.rpo_ro_attention_fwd(qkv.data_ptr(), qp.data_ptr(), sentinel.data_ptr(), out_p2.data_ptr(),
                                        off_d.data_ptr(), G, K, H, max(lens), causal, 0, code, st()))
    assert torch.equal(out_p2, out_p)
    assert torch.all(sentinel == 7.0)
    # backward: dq of the prompt queries vs autograd through the fp32 reference
    d_out = randn(G * K, D, dtype=dt, seed=42)
    dq = torch.zeros_like(qp)
    _lib.check(lib.rpo_ro_attention_bwd(qkv.data_ptr(), qp.data_ptr(), out_p.data_ptr(), d_out.data_ptr(), dq.data_ptr(),
                                        off_d.data_ptr(), G, K, H, max(lens), code, st()))
    qpr = qp.cpu().float().requires_grad_(True)
    _, rp2 = ref_attention(qkv.cpu(), qpr, off, K, H, causal, True)
    rp2.backward(d_out.cpu().float())
    assert relmax(dq.cpu(), qpr.grad) <= tol


DENSE_CASES = [
    # (name, G, n_ctx, K, H): the vision tower's shapes and the edges of what the tcgen05 kernel accepts
    ("vitb16_k24", 3, 197, 24, 12), ("vitb16_k4", 2, 197, 4, 12), ("vitb16_k48", 2, 197, 48, 12),
    ("one_tile", 2, 50, 24, 2), ("exact_128", 2, 128, 16, 2), ("n256", 2, 256, 100, 1), ("tiny", 1, 17, 5, 2),
    ("no_prompts", 2, 197, 0, 2), ("vitb16_k8", 2, 197, 8, 12), ("vitb16_k16", 2, 197, 16, 12),
    # several work items per CTA: K/V double buffer, Q ring and both TMEM slots wrap around (config 2 / 5 sizes)
    ("vitb16_k24_b32", 32, 197, 24, 12), ("vitb16_k48_b64", 64, 197, 48, 12), ("vitb16_ctx_only_b32", 32, 197, 0, 12),
    # ViT-L/14 (BASELINE config 3): 257 keys = two UMMA N blocks in one 512-column TMEM slot, three query tiles
    ("vitl14_k24", 2, 257, 24, 16), ("vitl14_k24_b16", 16, 257, 24, 16), ("n272", 3, 272, 0, 2), ("n16_one_block", 3, 9, 3, 1),
    # the widest two-slot shape (14 key blocks, 7 per softmax thread), an odd block count with prompts, one key block more
    # than ViT-B/16, and the first shape of the single-slot kernel
    ("n224_k16", 2, 224, 16, 2), ("n200_k40", 3, 200, 40, 3), ("n209_k8", 2, 209, 8, 2), ("n225_k0", 2, 225, 0, 2),
]


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_ro_attention_fwd_dense_tcgen05(prec, case):
    """The tcgen05 attention (vision form) against the fp32 reference and against the mma.sync kernel."""
    lib = _lib.load()
    _, G, n, K, H = case
    dt = DT[prec]
    D = H * 64
    off = [g * n for g in range(G + 1)]
    qkv = randn(G * n, 3 * D, dtype=dt, seed=60)
    qp = randn(max(1, G * K), D, dtype=dt, seed=61)
    out_ctx = torch.zeros(G * n, D, dtype=dt, device=dev())
    out_p = torch.full((max(1, G * K), D), 7.0, dtype=dt, device=dev())
    code = _lib.dtype_code(dt)
    _lib.check(lib.rpo_ro_attention_fwd_dense(qkv.data_ptr(), qp.data_ptr(), out_ctx.data_ptr(), out_p.data_ptr(), G, n,
                                              K, H, code, st()))
    torch.cuda.synchronize()
    rc, rp = ref_attention(qkv.cpu(), qp.cpu()[:G * K], off, K, H, 0, True)
    tol = TOL[prec] * 2
    assert relmax(out_ctx.cpu(), rc) <= tol
    if K:
        assert relmax(out_p.cpu()[:G * K], rp) <= tol
    # cross-check with the mma.sync path on identical inputs
    off_d = torch.tensor(off, dtype=torch.int32, device=dev())
    oc2 = torch.zeros_like(out_ctx)
    op2 = torch.zeros_like(out_p)
    _lib.check(lib.rpo_ro_attention_fwd(qkv.data_ptr(), qp.data_ptr(), oc2.data_ptr(), op2.data_ptr(), off_d.data_ptr(),
                                        G, K, H, n, 0, 1, code, st()))
    assert relmax(out_ctx, oc2) <= tol


def test_ro_attention_fwd_dense_rejects_unsupported():
    lib = _lib.load()
    x = torch.zeros(8, dtype=torch.float16, device=dev())
    # more than 288 context rows do not fit S + O into the 512 tensor-memory columns
    assert lib.rpo_ro_attention_fwd_dense(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 300, 24, 16,
                                          _lib.RPO_F16, st()) == -1
    assert lib.rpo_ro_attention_fwd_dense_supported(_lib.RPO_F16, 257, 24, 16) == 1   # ViT-L/14
    assert lib.rpo_ro_attention_fwd_dense_supported(_lib.RPO_F16, 300, 24, 16) == 0
    # prompts that would straddle two query tiles
    assert lib.rpo_ro_attention_fwd_dense(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 197, 64, 12,
                                          _lib.RPO_F16, st()) == -1
    # fp32 goes through the exact SIMT kernels
    assert lib.rpo_ro_attention_fwd_dense(x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 197, 24, 12,
                                          _lib.RPO_F32, st()) == -1


# ---------------------------------------------------------------------------------------------------
def ref_logits(img_feat, text_feat, logit_scale, K):
    i = img_feat.float()
    t = text_feat.float()
    i = i / i.norm(dim=-1, keepdim=True)
    t = t / t.norm(dim=-1, keepdim=True)
    return torch.einsum("bkd,ckd->bc", i, t) * logit_scale.exp() / K


@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("B,Cn,K,E", [(2, 2, 4, 512), (32, 100, 24, 512), (5, 1000, 3, 128),
                                      # edges of the tcgen05 pair kernels: full / odd batch, odd class counts (dl rows that
                                      # are not 16-byte aligned), one pair, the ViT-L/14 embedding width, a width the
                                      # tensor-core path does not take (falls back to the SIMT GEMM)
                                      (64, 37, 2, 64), (16, 100, 24, 768), (1, 3, 1, 128), (33, 1000, 2, 512),
                                      (7, 129, 3, 96)])
def test_logits_ce_fwd_bwd(prec, B, Cn, K, E):
    lib = _lib.load()
    dt = DT[prec]
    d = dev()
    img = randn(B, K, E, dtype=dt, seed=50)
    txt = randn(Cn, K, E, dtype=dt, seed=51)
    ls = torch.tensor(2.6592600369327783, device=d)
    label = (torch.arange(B) % Cn).to(d)
    img_n, img_s, text_n = torch.empty_like(img), torch.empty_like(img), torch.empty_like(txt)
    inorm = torch.empty(B * K + B, device=d)
    tnorm = torch.empty(Cn * K, device=d)
    pair = torch.empty(K, B, Cn, dtype=dt, device=d)
    logits = torch.empty(B, Cn, device=d)
    dlogits = torch.empty(B, Cn, device=d)
    loss = torch.zeros((), device=d)
    code = _lib.dtype_code(dt)
    _lib.check(lib.rpo_logits_ce_fwd(img.data_ptr(), txt.data_ptr(), ls.data_ptr(), label.data_ptr(), B, Cn, K, E,
                                     img_n.data_ptr(), img_s.data_ptr(), text_n.data_ptr(), inorm.data_ptr(),
                                     tnorm.data_ptr(), pair.data_ptr(), logits.data_ptr(), loss.data_ptr(),
                                     dlogits.data_ptr(), code, st()))
    ir = img.float().requires_grad_(True)
    tr = txt.float().requires_grad_(True)
    rl = ref_logits(ir, tr, ls, K)
    rloss = torch.nn.functional.cross_entropy(rl, label)
    rloss.backward()
    scale = float(ls.exp())
    tol = TOL[prec] * 3
    assert (logits - rl.detach()).abs().max().item() <= tol * scale
    assert abs(loss.item() - rloss.item()) <= tol * scale
    d_img, d_txt = torch.empty_like(img), torch.empty_like(txt)
    dl_t = torch.empty(B, Cn, dtype=dt, device=d)
    d_img_s, d_text_n = torch.empty_like(img), torch.empty_like(txt)
    _lib.check(lib.rpo_logits_ce_bwd(dlogits.data_ptr(), img.data_ptr(), txt.data_ptr(), img_n.data_ptr(),
                                     img_s.data_ptr(), text_n.data_ptr(), inorm.data_ptr(), tnorm.data_ptr(),
                                     ls.data_ptr(), B, Cn, K, E, dl_t.data_ptr(), d_img_s.data_ptr(),
                                     d_text_n.data_ptr(), d_img.data_ptr(), d_txt.data_ptr(), code, st()))
    gtol = 2e-2 if prec != "fp32" else 1e-4  # gradients of a 16-bit softmax: dominated by logit rounding
    assert relmax(d_img, ir.grad) <= gtol
    assert relmax(d_txt, tr.grad) <= gtol
    # eval mode: no label -> logits only
    logits2 = torch.empty(B, Cn, device=d)
    _lib.check(lib.rpo_logits_ce_fwd(img.data_ptr(), txt.data_ptr(), ls.data_ptr(), None, B, Cn, K, E,
                                     img_n.data_ptr(), img_s.data_ptr(), text_n.data_ptr(), inorm.data_ptr(),
                                     tnorm.data_ptr(), pair.data_ptr(), logits2.data_ptr(), None, None, code, st()))
    assert torch.equal(logits, logits2)


def test_sgd_step_matches_torch():
    lib = _lib.load()
    d = dev()
    for dt in (torch.float32, torch.float16):
        p = randn(24, 768, dtype=dt, seed=60)
        ref_p = torch.nn.Parameter(p.clone().float())
        opt = torch.optim.SGD([ref_p], lr=0.01, momentum=0.9, weight_decay=5e-4)
        buf = torch.zeros(p.numel(), device=d)
        lr = torch.tensor(0.01, device=d)
        first = torch.ones(1, dtype=torch.int32, device=d)
        mine = p.clone()
        for it in range(3):
            g = randn(24, 768, dtype=torch.float32, seed=61 + it)
            ref_p.grad = g.clone()
            opt.step()
            _lib.check(lib.rpo_sgd_step(mine.data_ptr(), _lib.dtype_code(dt), g.data_ptr(), buf.data_ptr(), p.numel(),
                                        lr.data_ptr(), 0.9, 5e-4, 1.0, first.data_ptr(), st()))
            first.zero_()
        tol = 1e-6 if dt == torch.float32 else 2e-3
        assert relmax(mine, ref_p.detach()) <= tol
