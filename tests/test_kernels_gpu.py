"""Kernel-level parity (GPU): every sm_100a kernel against the same op in plain torch fp32 on the same inputs.

All calls go through the C ABI (ctypes); torch is only the checker here.  Tolerances are written per test: the
kernels accumulate in fp32 and round once to fp16, so the bound is a few fp16 ulps of the output magnitude.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _pack_conv_w(w):  # OIHW -> [O][kh*kw][I] fp16
    O, I, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(O, kh * kw * I).contiguous().half()


def _close(got, ref, rtol, atol, what=""):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = (err > bound).sum().item()
    assert bad == 0, f"{what}: {bad} / {err.numel()} elements out of tolerance, max err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}"


# ------------------------------------------------------------------------------------------------ GEMM / linear
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 128, 128, 128), (1000, 320, 320, 0), (4096, 640, 1024, 0),
                                      (300, 1280, 640, 0), (512, 256, 512, 256), (130, 160, 192, 160)])
def test_linear(eng_mod, M, N, K, bn):
    x = _rand(1, M, K, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2).half()
    b = _rand(N, seed=3).float()
    out = torch.zeros(1, M, N, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, K, K)], w, N, out, B=1, Hin=1, Win=M, bias=b, out_ld=N, out_bstride=M * N, force_block_n=bn)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    _close(out, ref, 2e-3, 2e-3, f"linear {M}x{N}x{K}")


def test_linear_residual_batched_inplace(eng_mod):
    B, M, N, K = 3, 200, 320, 1280
    x = _rand(B, M, K, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2).half()
    b = _rand(N, seed=3).float()
    h = _rand(B, M, N, seed=4).half()
    ref = (x.float() @ w.float().t() + b).half().float() + h.float()
    eng_mod.k_conv_gemm([(x, K, K)], w, N, h, B=B, Hin=1, Win=M, bias=b, out_ld=N, out_bstride=M * N, res=(h, N, M * N))
    torch.cuda.synchronize()
    _close(h, ref, 2e-3, 2e-3, "linear+res in place")


def test_linear_transposed_store(eng_mod):
    B, M, N, K = 2, 328, 320, 1024
    x = _rand(B, M, K, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2).half()
    out = torch.zeros(B, N, M, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, K, K)], w, N, out, B=B, Hin=1, Win=M, mode=1, out_ld=M, out_bstride=N * M)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t()).transpose(1, 2)
    _close(out, ref, 2e-3, 2e-3, "V^T store")


def test_geglu(eng_mod):
    B, M, C = 2, 300, 320
    F4 = 4 * C
    x = _rand(B, M, C, seed=1).half()
    w = _rand(2 * F4, C, scale=C ** -0.5, seed=2).half()
    b = _rand(2 * F4, seed=3).float()
    # interleave rows in 128-blocks: [value block t | gate block t]
    idx = []
    for t in range(F4 // 128):
        idx += list(range(t * 128, t * 128 + 128)) + list(range(F4 + t * 128, F4 + t * 128 + 128))
    idx = torch.tensor(idx, device=DEV)
    wp, bp = w[idx].contiguous(), b[idx].contiguous()
    out = torch.zeros(B, M, F4, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, C, C)], wp, 2 * F4, out, B=B, Hin=1, Win=M, mode=2, bias=bp, out_ld=F4, out_bstride=M * F4)
    torch.cuda.synchronize()
    y = (x.float() @ w.float().t() + b).half()
    ref = y[..., :F4].float() * F.gelu(y[..., F4:].float()).half().float()
    _close(out, ref, 3e-3, 3e-3, "GEGLU")


def test_batched_scores_f32(eng_mod):
    B, L, D = 2, 320, 512
    q = _rand(B, L, D, seed=1).half()
    k = _rand(B, L, D, seed=2).half()
    out = torch.zeros(B, L, L, dtype=torch.float32, device=DEV)
    sc = 1.0 / math.sqrt(D)
    eng_mod.k_conv_gemm([(q, D, D)], k, L, out, B=B, Hin=1, Win=L, mode=3, scale=sc, w_bstride=L * D, out_ld=L, out_bstride=L * L)
    torch.cuda.synchronize()
    ref = torch.einsum("bld,bmd->blm", q.float(), k.float()) * sc
    _close(out, ref, 1e-4, 1e-3, "QK^T fp32")


# ------------------------------------------------------------------------------------------------ conv 3x3
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 16, 16, 64, 64), (2, 32, 32, 128, 128), (1, 128, 128, 128, 256), (2, 40, 40, 320, 320),
                                            (1, 10, 10, 1280, 1280), (1, 8, 8, 64, 160), (1, 4, 4, 128, 64), (1, 2, 2, 64, 64), (1, 1, 1, 64, 64)])
def test_conv3x3(eng_mod, B, H, W, Cin, Cout):
    x = _rand(B, H, W, Cin, seed=1).half()
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, Cin, Cin)], _pack_conv_w(w), Cout, out, B=B, Hin=H, Win=W, ksize=3, bias=b, out_ld=Cout,
                        out_bstride=H * W * Cout)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, f"conv3x3 {B}x{H}x{W} {Cin}->{Cout}")


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("H,W,C", [(32, 32, 128), (16, 16, 320), (8, 8, 64), (2, 2, 64)])
def test_conv3x3_stride2(eng_mod, pad, H, W, C):
    B = 2
    x = _rand(B, H, W, C, seed=1).half()
    w = _rand(C, C, 3, 3, scale=(9 * C) ** -0.5, seed=2).half()
    b = _rand(C, seed=3).float()
    out = torch.zeros(B, H // 2, W // 2, C, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, C, C)], _pack_conv_w(w), C, out, B=B, Hin=H, Win=W, ksize=3, stride=2, pad=pad, bias=b, out_ld=C,
                        out_bstride=(H // 2) * (W // 2) * C)
    torch.cuda.synchronize()
    xin = x.float().permute(0, 3, 1, 2)
    if pad == 0:
        ref = F.conv2d(xin, w.float(), b, stride=2, padding=1)
    else:  # VAE Downsample2D: F.pad (0,1,0,1) then stride-2 conv without padding
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w.float(), b, stride=2, padding=0)
    _close(out, ref.permute(0, 2, 3, 1), 2e-3, 2e-3, "conv s2")


def test_conv1x1_concat_two_sources(eng_mod):
    B, H, W, C0, C1, Cout = 2, 16, 16, 640, 320, 640
    a = _rand(B, H, W, C0, seed=1).half()
    s = _rand(B, H, W, C1, seed=2).half()
    w = _rand(Cout, C0 + C1, 1, 1, scale=(C0 + C1) ** -0.5, seed=3).half()
    b = _rand(Cout, seed=4).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(a, C0, C0), (s, C1, C1)], _pack_conv_w(w), Cout, out, B=B, Hin=H, Win=W, ksize=1, bias=b, out_ld=Cout,
                        out_bstride=H * W * Cout)
    torch.cuda.synchronize()
    ref = F.conv2d(torch.cat([a, s], -1).float().permute(0, 3, 1, 2), w.float(), b).permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, "1x1 concat")


def test_conv3x3_residual_upsample_store(eng_mod):
    B, H, W, C = 1, 16, 16, 128
    x = _rand(B, H, W, C, seed=1).half()
    r = _rand(B, H, W, C, seed=5).half()
    w = _rand(C, C, 3, 3, scale=(9 * C) ** -0.5, seed=2).half()
    b = _rand(C, seed=3).float()
    out = torch.zeros(B, 2 * H, 2 * W, C, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, C, C)], _pack_conv_w(w), C, out, B=B, Hin=H, Win=W, ksize=3, bias=b, ups2=1, out_ld=C,
                        out_bstride=4 * H * W * C, res=(r, C, H * W * C))
    torch.cuda.synchronize()
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).half().float() + r.float().permute(0, 3, 1, 2)
    ref = F.interpolate(y, scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, "conv+res+ups2")


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, heads, bias, scale):
    B, Lq, C = q.shape
    Lk = k.shape[1]
    qh = q.float().view(B, Lq, heads, 64).transpose(1, 2)
    kh = k.float().view(B, Lk, heads, 64).transpose(1, 2)
    vh = v.float().view(B, Lk, heads, 64).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if bias is not None:
        s = s + bias[:, None, None, :Lk]
    p = s.softmax(-1)
    return torch.matmul(p, vh).transpose(1, 2).reshape(B, Lq, C)


@pytest.mark.parametrize("B,heads,Lq,Lk,with_bias", [(1, 1, 256, 256, False), (2, 5, 1024, 1024, True), (1, 10, 400, 1600, False),
                                                     (1, 20, 64, 64, True), (1, 2, 100, 4096, False), (2, 5, 1600, 1600, True),
                                                     (1, 1, 128, 128, True)])
def test_attention(eng_mod, B, heads, Lq, Lk, with_bias):
    C = heads * 64
    q = _rand(B, Lq, C, seed=1).half()
    k = _rand(B, Lk, C, seed=2).half()
    v = _rand(B, Lk, C, seed=3).half()
    ldvt = (Lk + 7) // 8 * 8
    vt = torch.zeros(B, C, ldvt, dtype=torch.float16, device=DEV)
    vt[:, :, :Lk] = v.transpose(1, 2)
    bias = None
    lpad = (Lk + 127) // 128 * 128
    if with_bias:
        g = torch.Generator().manual_seed(7)
        lv = torch.randint(0, 3, (B, Lk), generator=g).float()  # 0: fg, 1: unknown, 2: bg
        lv[:, 0] = 0  # at least one foreground key per sample
        bias = torch.full((B, lpad), float("-inf"))
        bias[:, :Lk] = lv * -5000.0
        bias = bias.to(DEV)
    out = torch.zeros(B, Lq, C, dtype=torch.float16, device=DEV)
    # the kernel takes the additive key bias pre-multiplied by log2(e)
    bias_l2 = None if bias is None else (bias * math.log2(math.e)).contiguous()
    eng_mod.k_attention(q, k, vt, out, B=B, heads=heads, Lq=Lq, Lk=Lk, ldq=C, ldk=C, ldvt=ldvt, ldo=C, bias=bias_l2, bias_bstride=lpad)
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, heads, bias, 0.125)
    _close(out, ref, 4e-3, 2e-3, f"attention B{B} h{heads} {Lq}x{Lk} bias={with_bias}")


def test_attention_soft_bias_no_foreground(eng_mod):
    """All keys unknown/background (bias -5000/-10000): softmax renormalises over the least-negative keys."""
    B, heads, L = 1, 2, 256
    C = heads * 64
    q, k, v = _rand(B, L, C, seed=1).half(), _rand(B, L, C, seed=2).half(), _rand(B, L, C, seed=3).half()
    vt = v.transpose(1, 2).contiguous()
    bias = torch.full((B, L), -10000.0)
    bias[:, ::3] = -5000.0
    bias = bias.to(DEV)
    out = torch.zeros(B, L, C, dtype=torch.float16, device=DEV)
    eng_mod.k_attention(q, k, vt, out, B=B, heads=heads, Lq=L, Lk=L, ldq=C, ldk=C, ldvt=L, ldo=C, bias=(bias * math.log2(math.e)).contiguous(), bias_bstride=L)
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, heads, bias, 0.125)
    _close(out, ref, 6e-3, 3e-3, "attention soft bias")


# ------------------------------------------------------------------------------------------------ norms / softmax
@pytest.mark.parametrize("B,HW,C,silu,eps", [(2, 4096, 128, 1, 1e-6), (1, 1600, 320, 1, 1e-5), (2, 256, 1280, 0, 1e-6), (1, 64, 512, 1, 1e-6),
                                             (3, 1, 320, 1, 1e-5)])
def test_groupnorm(eng_mod, B, HW, C, silu, eps):
    x = (_rand(B, HW, C, seed=1) * 2 + 0.5).half()
    g = _rand(C, seed=2).float() * 0.2 + 1.0
    b = _rand(C, seed=3).float() * 0.2
    out = torch.zeros(B, HW, C, dtype=torch.float16, device=DEV)
    eng_mod.k_groupnorm([(x, C, C)], g, b, out, B=B, HW=HW, eps=eps, silu=silu)
    torch.cuda.synchronize()
    ref = F.group_norm(x.float().transpose(1, 2), 32, g, b, eps)
    if silu:
        ref = F.silu(ref)
    _close(out, ref.transpose(1, 2), 2e-3, 2e-3, "groupnorm")


def test_groupnorm_concat_straddling_groups(eng_mod):
    B, HW, C0, C1 = 2, 1024, 1280, 640  # 1920 channels: 60 per group, groups straddle the source boundary
    a = _rand(B, HW, C0, seed=1).half()
    s = (_rand(B, HW, C1, seed=2) * 3).half()
    g = _rand(C0 + C1, seed=3).float() * 0.2 + 1.0
    b = _rand(C0 + C1, seed=4).float() * 0.2
    out = torch.zeros(B, HW, C0 + C1, dtype=torch.float16, device=DEV)
    eng_mod.k_groupnorm([(a, C0, C0), (s, C1, C1)], g, b, out, B=B, HW=HW, eps=1e-5, silu=1)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(torch.cat([a, s], -1).float().transpose(1, 2), 32, g, b, 1e-5)).transpose(1, 2)
    _close(out, ref, 2e-3, 2e-3, "groupnorm concat")


@pytest.mark.parametrize("rows,C", [(1000, 320), (512, 640), (77, 1280)])
def test_layernorm(eng_mod, rows, C):
    x = (_rand(rows, C, seed=1) * 1.5 + 0.3).half()
    g = _rand(C, seed=2).float() * 0.2 + 1.0
    b = _rand(C, seed=3).float() * 0.2
    y = torch.zeros_like(x)
    eng_mod.k_layernorm(x, y, g, b, rows, C)
    torch.cuda.synchronize()
    _close(y, F.layer_norm(x.float(), (C,), g, b, 1e-5), 2e-3, 2e-3, "layernorm")


@pytest.mark.parametrize("rows,L", [(64, 4096), (7, 16384), (33, 1024), (5, 64)])
def test_softmax_rows(eng_mod, rows, L):
    s = _rand(rows, L, seed=1).float() * 4
    p = torch.zeros(rows, L, dtype=torch.float16, device=DEV)
    eng_mod.k_softmax_rows(s, p, rows, L)
    torch.cuda.synchronize()
    _close(p, s.softmax(-1), 2e-3, 1e-6, "softmax rows")


# ------------------------------------------------------------------------------------------------ small convs
@pytest.mark.parametrize("Cin,Cout,k", [(4, 128, 3), (8, 320, 3), (4, 1024, 3), (8, 8, 1)])
def test_small_cin_conv(eng_mod, Cin, Cout, k):
    B, H, W = 2, 24, 24
    x = _rand(B, H, W, Cin, seed=1).half()
    w = _rand(Cout, Cin, k, k, scale=(k * k * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    eng_mod.k_direct_conv(x, _pack_conv_w(w), b, out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=k, x_ld=Cin, out_ld=Cout)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=k // 2).permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, "small-cin conv")


@pytest.mark.parametrize("Cin,Cout", [(320, 4), (512, 8)])
def test_small_cout_conv(eng_mod, Cin, Cout):
    B, H, W = 2, 16, 16
    x = _rand(B, H, W, Cin, seed=1).half()
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    eng_mod.k_direct_conv(x, _pack_conv_w(w), b, out, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=3, x_ld=Cin, out_ld=Cout)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, "small-cout conv")


def test_attention_late_rescale(eng_mod):
    """Keys whose scores grow along the sequence force the lazy-rescale path (O rows rescaled in TMEM) on late tiles."""
    B, heads, Lq, Lk = 1, 2, 256, 1024
    C = heads * 64
    q = _rand(B, Lq, C, seed=1).half()
    k = _rand(B, Lk, C, seed=2)
    ramp = torch.linspace(0.2, 3.0, Lk, device=DEV)[None, :, None]
    k = (k * ramp).half()
    v = _rand(B, Lk, C, seed=3).half()
    vt = v.transpose(1, 2).contiguous()
    out = torch.zeros(B, Lq, C, dtype=torch.float16, device=DEV)
    eng_mod.k_attention(q, k, vt, out, B=B, heads=heads, Lq=Lq, Lk=Lk, ldq=C, ldk=C, ldvt=Lk, ldo=C)
    torch.cuda.synchronize()
    _close(out, _attn_ref(q, k, v, heads, None, 0.125), 4e-3, 2e-3, "attention late rescale")


@pytest.mark.parametrize("start", [300, 330, 360, 384])  # first foreground key in chunk 1 / 2 / 3 / 0 of its 128-key tile
def test_attention_bias_background_first(eng_mod, start):
    """First key tiles are all background (-10000), foreground keys only appear later: the reference max jumps by ~14 000
    (log2 units) and everything accumulated so far must vanish (redo path of the softmax, before and after the low half
    of P has been handed to the tensor core)."""
    B, heads, L = 1, 1, 512
    C = 64
    q, k, v = _rand(B, L, C, seed=1).half(), _rand(B, L, C, seed=2).half(), _rand(B, L, C, seed=3).half()
    vt = v.transpose(1, 2).contiguous()
    bias = torch.full((B, L), -10000.0)
    bias[:, start:start + 40] = 0.0
    bias = bias.to(DEV)
    out = torch.zeros(B, L, C, dtype=torch.float16, device=DEV)
    eng_mod.k_attention(q, k, vt, out, B=B, heads=heads, Lq=L, Lk=L, ldq=C, ldk=C, ldvt=L, ldo=C, bias=(bias * math.log2(math.e)).contiguous(), bias_bstride=L)
    torch.cuda.synchronize()
    _close(out, _attn_ref(q, k, v, heads, bias, 0.125), 4e-3, 2e-3, "attention bg-first")


def test_skinny_cout_conv_tensor_path(eng_mod):
    """3x3 conv with 4 output channels on the tcgen05 path (BLOCK_N=16), narrow store + fp16 division (UNet conv_out / 0.18215)."""
    B, H, W, Cin = 2, 16, 16, 320
    x = _rand(B, H, W, Cin, seed=1).half()
    w = _rand(4, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(4, seed=3).float()
    wp = torch.zeros(8, 9 * Cin, dtype=torch.float16, device=DEV)
    wp[:4] = _pack_conv_w(w)
    bp = torch.zeros(8, device=DEV)
    bp[:4] = b
    out = torch.zeros(B, H, W, 4, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, Cin, Cin)], wp, 8, out, B=B, Hin=H, Win=W, ksize=3, bias=bp, out_ld=4, out_bstride=H * W * 4, n_store=4,
                        post_div=0.18215)
    torch.cuda.synchronize()
    ref = (F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).half().float() / 0.18215).permute(0, 2, 3, 1)
    _close(out, ref, 2e-3, 2e-3, "skinny conv")


def test_alpha_head_epilogue(eng_mod):
    B, H, W, Cin = 1, 32, 32, 128
    x = _rand(B, H, W, Cin, seed=1).half()
    w = _rand(3, Cin, 3, 3, scale=(9 * Cin) ** -0.5 * 1.5, seed=2).half()
    b = _rand(3, seed=3).float() * 0.1
    wp = torch.zeros(8, 9 * Cin, dtype=torch.float16, device=DEV)
    wp[:3] = _pack_conv_w(w)
    bp = torch.zeros(8, device=DEV)
    bp[:3] = b
    alpha = torch.zeros(B, H, W, dtype=torch.float16, device=DEV)
    pre = torch.zeros(B, H, W, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, Cin, Cin)], wp, 8, alpha, B=B, Hin=H, Win=W, ksize=3, bias=bp, mode=4, out_ld=1, out_bstride=H * W, out2=pre)
    torch.cuda.synchronize()
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).half()
    m = y.float().mean(1).half()
    ref = ((m.float().clamp(-1, 1) + 1).half().float() / 2).half()
    _close(pre, m, 2e-3, 2e-3, "alpha head pre-clip mean")
    _close(alpha, ref, 2e-3, 1e-3, "alpha head")


def test_silu_accuracy_via_groupnorm(eng_mod):
    """The one-MUFU SiLU (bit-trick reciprocal + 2 Newton steps) stays within 1 fp16 ulp of torch's SiLU."""
    B, HW, C = 1, 4096, 128
    x = (_rand(B, HW, C, seed=1) * 4).half()
    g = torch.ones(C, device=DEV) * 3.0
    b = torch.zeros(C, device=DEV)
    out = torch.zeros(B, HW, C, dtype=torch.float16, device=DEV)
    eng_mod.k_groupnorm([(x, C, C)], g, b, out, B=B, HW=HW, eps=1e-6, silu=1)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(x.float().transpose(1, 2), 32, g, b, 1e-6)).transpose(1, 2)
    _close(out, ref, 1.2e-3, 2e-4, "silu")


@pytest.mark.parametrize("B,H,W,Cin", [(3, 16, 8, 128), (2, 32, 32, 128), (1, 128, 128, 64), (5, 8, 16, 192)])
def test_conv3x3_two_m_subtiles(eng_mod, B, H, W, Cin):
    """256x128 CTA tile (two M sub-tiles share one weight tile), incl. an odd number of M tiles and a residual."""
    Cout = 128
    x = _rand(B, H, W, Cin, seed=1).half()
    r = _rand(B, H, W, Cout, seed=5).half()
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, Cin, Cin)], _pack_conv_w(w), Cout, out, B=B, Hin=H, Win=W, ksize=3, bias=b, out_ld=Cout,
                        out_bstride=H * W * Cout, res=(r, Cout, H * W * Cout), force_mt=2)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).half().float().permute(0, 2, 3, 1) + r.float()
    _close(out, ref, 2e-3, 2e-3, "conv3x3 MT=2")


def test_linear_two_m_subtiles(eng_mod):
    B, M, N, K = 2, 700, 640, 320
    x = _rand(B, M, K, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2).half()
    b = _rand(N, seed=3).float()
    out = torch.zeros(B, M, N, dtype=torch.float16, device=DEV)
    eng_mod.k_conv_gemm([(x, K, K)], w, N, out, B=B, Hin=1, Win=M, bias=b, out_ld=N, out_bstride=M * N, force_block_n=128, force_mt=2)
    torch.cuda.synchronize()
    _close(out, x.float() @ w.float().t() + b, 2e-3, 2e-3, "linear MT=2")


@pytest.mark.parametrize("B,H,W,Cin,Cout,mt", [(2, 32, 32, 128, 128, 2), (1, 40, 40, 320, 320, 0), (2, 16, 16, 256, 256, 0), (3, 8, 8, 64, 64, 0),
                                               (2, 256, 256, 64, 128, 2)])  # last: 512 slots -> coalesced two-stage reduction of the partials
def test_conv_epilogue_groupnorm_partials(eng_mod, B, H, W, Cin, Cout, mt):
    """The conv epilogue's per-(tile, channel) sums of the STORED fp16 outputs reproduce the tensor's channel statistics,
    and GroupNorm fed with them equals GroupNorm computing its own statistics (bit for bit run to run)."""
    x = _rand(B, H, W, Cin, seed=1).half()
    r = _rand(B, H, W, Cout, seed=5).half()
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    slots = eng_mod.conv_tiles_per_image(H, W)
    stats = torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
    eng_mod.k_conv_gemm([(x, Cin, Cin)], _pack_conv_w(w), Cout, out, B=B, Hin=H, Win=W, ksize=3, bias=b, out_ld=Cout, out_bstride=H * W * Cout,
                        res=(r, Cout, H * W * Cout), force_mt=mt, stats=stats)
    torch.cuda.synchronize()
    assert torch.isfinite(stats).all()
    tot = stats.double().sum(1)  # (B, C, 2)
    o = out.double().view(B, H * W, Cout)
    assert torch.allclose(tot[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)
    g = _rand(Cout, seed=7).float() * 0.2 + 1.0
    bt = _rand(Cout, seed=8).float() * 0.2
    y1 = torch.zeros_like(out).view(B, H * W, Cout)
    y2 = torch.zeros_like(y1)
    eng_mod.k_groupnorm([(out, Cout, Cout)], g, bt, y1, B=B, HW=H * W, eps=1e-5, silu=1)
    eng_mod.k_groupnorm([(out, Cout, Cout)], g, bt, y2, B=B, HW=H * W, eps=1e-5, silu=1, pre=[stats], pre_slots=slots)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(out.float().view(B, H * W, Cout).transpose(1, 2), 32, g, bt, 1e-5)).transpose(1, 2)
    _close(y2, ref, 2e-3, 2e-3, "groupnorm from epilogue partials")
    assert (y1.float() - y2.float()).abs().max().item() <= 2e-3


def test_groupnorm_two_sources_many_partial_slots(eng_mod):
    """concat GroupNorm fed with per-slot partials of BOTH sources (1024 slots each -> the pre-reduction kernel runs per
    source); must equal the GroupNorm that computes its own statistics, and be bit-identical for a sample alone vs in a batch."""
    B, HW, C0, C1, slots = 2, 65536, 128, 64, 1024
    a = (_rand(B, HW, C0, seed=1) + 0.3).half()
    s2 = (_rand(B, HW, C1, seed=2) * 2).half()
    g = _rand(C0 + C1, seed=3).float() * 0.2 + 1.0
    bt = _rand(C0 + C1, seed=4).float() * 0.2

    def partials(t, C):
        v = t.float().view(B, slots, HW // slots, C)
        return torch.stack([v.sum(2), (v * v).sum(2)], dim=-1).contiguous()  # (B, slots, C, 2)

    pa, ps = partials(a, C0), partials(s2, C1)
    y1 = torch.zeros(B, HW, C0 + C1, dtype=torch.float16, device=DEV)
    y2 = torch.zeros_like(y1)
    eng_mod.k_groupnorm([(a, C0, C0), (s2, C1, C1)], g, bt, y1, B=B, HW=HW, eps=1e-5, silu=1)
    eng_mod.k_groupnorm([(a, C0, C0), (s2, C1, C1)], g, bt, y2, B=B, HW=HW, eps=1e-5, silu=1, pre=[pa, ps], pre_slots=slots)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(torch.cat([a, s2], -1).float().transpose(1, 2), 32, g, bt, 1e-5)).transpose(1, 2)
    _close(y2, ref, 2e-3, 2e-3, "groupnorm two-source partials")
    assert (y1.float() - y2.float()).abs().max().item() <= 2e-3
    y3 = torch.zeros(1, HW, C0 + C1, dtype=torch.float16, device=DEV)
    eng_mod.k_groupnorm([(a[1:], C0, C0), (s2[1:], C1, C1)], g, bt, y3, B=1, HW=HW, eps=1e-5, silu=1, pre=[pa[1:].contiguous(), ps[1:].contiguous()], pre_slots=slots)
    torch.cuda.synchronize()
    assert torch.equal(y3[0], y2[1])


# ------------------------------------------------------------------------------------------------ key compaction (attn1)
def _compact_ref(bias_l2, L):
    """torch restatement of key_compact_kernel for one sample: (idx, cbias, ntiles)."""
    b = bias_l2[:L]
    keep = (b >= b.max() - 2500.0 * math.log2(math.e)).nonzero().flatten()
    n = keep.numel()
    padded = (n + 127) // 128 * 128
    idx = torch.cat([keep, keep[:1].repeat(padded - n)])
    cb = torch.cat([b[keep], torch.full((padded - n,), float("-inf"), device=b.device)])
    return idx, cb, padded // 128


@pytest.mark.parametrize("B,L", [(3, 1024), (2, 16384), (4, 64), (1, 4096)])
def test_key_compact_and_gather(eng_mod, B, L):
    lpad = (L + 127) // 128 * 128
    g = torch.Generator().manual_seed(L + B)
    lv = torch.randint(0, 3, (B, L), generator=g).float()
    if B > 1:
        lv[1] = lv[1].clamp(min=1)  # sample 1 has no foreground key: unknown keys (-5000) are the maximum
    if B > 2:
        lv[2] = 2                   # all background: nothing can be dropped
    bias = torch.full((B, lpad), float("-inf"))
    bias[:, :L] = lv * -5000.0 * math.log2(math.e)
    bias = bias.to(DEV)
    cb = torch.zeros(B, lpad, device=DEV)
    idx = torch.full((B, lpad), -1, dtype=torch.int32, device=DEV)
    nt = torch.zeros(B, dtype=torch.int32, device=DEV)
    eng_mod.k_key_compact(bias, cb, idx, nt, B=B, L=L, lpad=lpad)
    C = 192
    src = _rand(B, L, C, seed=5).half()
    dst = torch.zeros_like(src)
    eng_mod.k_gather_rows(src, dst, idx, nt, B=B, L=L, C_=C, idx_bstride=lpad)
    torch.cuda.synchronize()
    for b in range(B):
        ri, rc, rn = _compact_ref(bias[b], L)
        assert int(nt[b]) == rn, (b, int(nt[b]), rn)
        n = rn * 128
        assert torch.equal(idx[b, :n].long(), ri), f"idx sample {b}"
        assert torch.equal(cb[b, :n], rc), f"cbias sample {b}"
        m = min(n, L)  # dst holds L rows; a padded count beyond L (L < 128) is covered by the TMA zero fill of the K/V tiles
        assert torch.equal(dst[b, :m], src[b, ri[:m]]), f"gather sample {b}"
        assert torch.count_nonzero(dst[b, m:]) == 0  # rows beyond the padded count are left alone


@pytest.mark.parametrize("B,heads,L", [(3, 2, 1024), (2, 5, 4096)])
def test_attention_compacted_keys_matches_full(eng_mod, B, heads, L):
    """attn1 with the dropped keys (probability exactly 0 under the -5000/-10000 bias) == the reference softmax over ALL keys;
    per-sample tile counts differ."""
    C = heads * 64
    q, k, v = _rand(B, L, C, seed=1).half(), _rand(B, L, C, seed=2).half(), _rand(B, L, C, seed=3).half()
    g = torch.Generator().manual_seed(11)
    lv = torch.randint(0, 3, (B, L), generator=g).float()
    lv[0, L // 3:] = 2          # sample 0: foreground only in the first third
    lv[1] = lv[1].clamp(min=1)  # sample 1: no foreground at all
    bias = (lv * -5000.0).to(DEV)
    bias_l2 = (bias * math.log2(math.e)).contiguous()
    cb = torch.zeros(B, L, device=DEV)
    idx = torch.zeros(B, L, dtype=torch.int32, device=DEV)
    nt = torch.zeros(B, dtype=torch.int32, device=DEV)
    eng_mod.k_key_compact(bias_l2, cb, idx, nt, B=B, L=L, lpad=L)
    kc, vc = torch.full_like(k, float("nan")), torch.full_like(v, float("nan"))  # rows beyond the kept keys must never be read
    eng_mod.k_gather_rows(k, kc, idx, nt, B=B, L=L, C_=C, idx_bstride=L)
    eng_mod.k_gather_rows(v, vc, idx, nt, B=B, L=L, C_=C, idx_bstride=L)
    vt = vc.transpose(1, 2).contiguous()
    out = torch.zeros(B, L, C, dtype=torch.float16, device=DEV)
    eng_mod.k_attention(q, kc, vt, out, B=B, heads=heads, Lq=L, Lk=L, ldq=C, ldk=C, ldvt=L, ldo=C, bias=cb, bias_bstride=L, ntiles=nt)
    torch.cuda.synchronize()
    assert int(nt[0]) < int(nt[1]) <= L // 128
    _close(out, _attn_ref(q, k, v, heads, bias, 0.125), 4e-3, 2e-3, "attention over compacted keys")


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 64, 128), (3, 32, 48, 256, 128)])
def test_conv1x1_two_m_subtiles_dual_epilogue(eng_mod, B, H, W, Cin, Cout):
    """1x1 conv with 1 / 4 K steps on 256x128 tiles: the two M sub-tiles are drained by two epilogue warpgroups (the im2col
    conv_in and the VAE shortcut convs); output and GroupNorm partials must match, odd number of M tiles included."""
    x = _rand(B, H, W, Cin, seed=1).half()
    w = _rand(Cout, Cin, scale=Cin ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    slots = eng_mod.conv_tiles_per_image(H, W)
    stats = torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
    eng_mod.k_conv_gemm([(x, Cin, Cin)], w, Cout, out, B=B, Hin=H, Win=W, ksize=1, bias=b, out_ld=Cout, out_bstride=H * W * Cout,
                        force_mt=2, stats=stats)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t() + b)
    _close(out, ref, 2e-3, 2e-3, "conv1x1 MT=2 EWG=2")
    assert torch.isfinite(stats).all()
    tot = stats.double().sum(1)
    o = out.double().view(B, H * W, Cout)
    assert torch.allclose(tot[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


# ------------------------------------------------------------------------------------------------ resident-halo 3x3 convs
@pytest.mark.parametrize("B,H,W,C0,C1,Cout,bn,mt,res,ups2", [
    (2, 32, 32, 128, 0, 128, 128, 2, True, 0),    # 256x128 tiles, residual K steps, GroupNorm partials
    (3, 16, 8, 64, 0, 128, 128, 2, False, 0),     # one patch per image, odd number of M tiles (past-the-end sub-tile)
    (1, 64, 64, 64, 0, 256, 256, 1, True, 0),     # 256-wide tiles
    (2, 16, 16, 320, 0, 320, 160, 1, True, 0),    # 160-wide tiles, two epilogue warpgroups
    (1, 32, 32, 192, 128, 256, 256, 1, False, 0),  # cat([h, skip]) input: two sources
    (1, 40, 40, 128, 0, 512, 256, 1, True, 0),    # ragged image (40 = 5 x 8 = 2.5 x 16): halo and tile rows outside the image, two N tiles
    (1, 16, 16, 128, 0, 256, 256, 1, True, 1),    # fused nearest-2x scatter store
    (2, 24, 24, 640, 0, 640, 160, 1, False, 1),
])
def test_conv3x3_halo(eng_mod, B, H, W, C0, C1, Cout, bn, mt, res, ups2):
    """3x3 conv issued from ONE resident (8+2) x (16+2) halo tile per 64-channel slice (nine shifted descriptor windows) against
    the reference and against the tap-per-TMA-box kernel (different K order: equal up to fp32 summation order)."""
    Cin = C0 + C1
    a = _rand(B, H, W, C0, seed=1).half()
    s2 = _rand(B, H, W, C1, seed=6).half() if C1 else None
    r = _rand(B, H, W, Cout, seed=5).half() if res else None
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    srcs = [(a, C0, C0)] + ([(s2, C1, C1)] if C1 else [])
    slots = eng_mod.conv_tiles_per_image(H, W)
    sc = 2 if ups2 else 1
    outs, stats = [], []
    for fh in (1, -1):
        out = torch.zeros(B, sc * H, sc * W, Cout, dtype=torch.float16, device=DEV)
        st = None if ups2 else torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
        eng_mod.k_conv_gemm(srcs, _pack_conv_w(w), Cout, out, B=B, Hin=H, Win=W, ksize=3, bias=b, ups2=ups2, out_ld=Cout,
                            out_bstride=sc * sc * H * W * Cout, res=(r, Cout, H * W * Cout) if res else None, force_block_n=bn,
                            force_mt=mt, stats=st, force_halo=fh, force_swap=-1)
        torch.cuda.synchronize()
        outs.append(out)
        stats.append(st)
    xin = torch.cat([a, s2], -1) if C1 else a
    y = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), b, padding=1).half().float()
    if res:
        y = y + r.float().permute(0, 3, 1, 2)
    if ups2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    ref = y.permute(0, 2, 3, 1)
    _close(outs[0], ref, 2e-3, 2e-3, "conv3x3 halo")
    d = (outs[0].float() - outs[1].float()).abs()
    assert d.max().item() <= 4e-3 and (d > 0).float().mean().item() < 0.05, f"halo vs tap kernels: max {d.max():.3e}, differing {(d > 0).float().mean():.4f}"
    if not ups2:
        assert torch.isfinite(stats[0]).all()
        tot = stats[0].double().sum(1)
        o = outs[0].double().view(B, H * W, Cout)
        assert torch.allclose(tot[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
        assert torch.allclose(tot[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


# ------------------------------------------------------------------------------------------------ swapped-operand 3x3 convs
@pytest.mark.parametrize("B,H,W,C0,C1,Cout,res", [
    (2, 32, 32, 128, 0, 128, True),     # the VAE case: 128 -> 128 with residual and GroupNorm partials
    (1, 64, 64, 256, 0, 128, False),    # 256 -> 128 (decoder conv1 of the last block)
    (3, 16, 16, 64, 0, 128, True),      # one patch per image
    (1, 48, 80, 128, 64, 256, True),    # two sources, two 128-channel N tiles
    (1, 40, 24, 64, 0, 128, True),      # ragged image: patch rows / columns outside the image
])
def test_conv3x3_swapped_operands(eng_mod, B, H, W, C0, C1, Cout, res):
    """3x3 conv computed as D^T = W . X^T (channels on the MMA's M, 256 pixels on N; conv_swap.cu): output, residual K steps and
    GroupNorm partials against the reference and the pixels-on-M kernel."""
    Cin = C0 + C1
    a = _rand(B, H, W, C0, seed=1).half()
    s2 = _rand(B, H, W, C1, seed=6).half() if C1 else None
    r = _rand(B, H, W, Cout, seed=5).half() if res else None
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    b = _rand(Cout, seed=3).float()
    srcs = [(a, C0, C0)] + ([(s2, C1, C1)] if C1 else [])
    slots = eng_mod.conv_tiles_per_image(H, W)
    if 2 * ((W + 15) // 16) * ((H + 15) // 16) != slots:
        pytest.skip("default patch gives a different GroupNorm slot count: the engine would not select the swapped kernel here")
    outs, stats = [], []
    for fs in (1, -1):
        out = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
        st = torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
        eng_mod.k_conv_gemm(srcs, _pack_conv_w(w), Cout, out, B=B, Hin=H, Win=W, ksize=3, bias=b, out_ld=Cout, out_bstride=H * W * Cout,
                            res=(r, Cout, H * W * Cout) if res else None, stats=st, force_swap=fs, force_halo=-1)
        torch.cuda.synchronize()
        outs.append(out)
        stats.append(st)
    xin = torch.cat([a, s2], -1) if C1 else a
    y = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), b, padding=1).half().float()
    if res:
        y = y + r.float().permute(0, 3, 1, 2)
    ref = y.permute(0, 2, 3, 1)
    _close(outs[0], ref, 2e-3, 2e-3, "conv3x3 swapped operands")
    d = (outs[0].float() - outs[1].float()).abs()
    assert d.max().item() <= 4e-3 and (d > 0).float().mean().item() < 0.05
    assert torch.isfinite(stats[0]).all()
    tot = stats[0].double().sum(1)
    o = outs[0].double().view(B, H * W, Cout)
    assert torch.allclose(tot[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


# ------------------------------------------------------------------------------------------------ GroupNorm fused into the consuming conv
@pytest.mark.parametrize("B,H,W,C0,C1,Cout,res,silu", [
    (2, 64, 64, 128, 0, 128, False, 1),   # VAE 128 -> 128 resnet conv1
    (1, 64, 64, 128, 0, 128, True, 1),    # conv2: + residual K steps (the residual boxes must pass the transform warps untouched)
    (1, 96, 48, 128, 64, 128, True, 1),   # concat input: groups straddle the two sources
    (1, 32, 64, 256, 0, 256, False, 0),   # two N tiles, no SiLU
    (2, 128, 128, 64, 0, 128, False, 1),  # many tiles per CTA: slot ring wraps, per-tile validity masks change
    # >= 74 (pixel tile, 256-channel) items: the CTA-PAIR form (cta_group::2, half a patch per CTA) is selected
    (2, 128, 96, 256, 0, 256, False, 0),  # pair, no SiLU: bit for bit against the single-CTA unfused path
    (4, 96, 64, 128, 0, 256, True, 0),    # pair + residual (four identity K slices, half-patch residual boxes), bit for bit
    (1, 160, 96, 256, 128, 512, True, 1), # pair, two channel-tile pairs, concat input, residual, SiLU
    (1, 224, 64, 512, 0, 512, False, 1),  # pair, eight input slices per tile: rings wrap many times inside one tile
])
def test_conv3x3_fused_groupnorm(eng_mod, B, H, W, C0, C1, Cout, res, silu):
    """GroupNorm(32)(+SiLU) applied to the conv's resident input halo tile by the transform warps of conv_swap_halo_kernel<true>
    against the stand-alone apply pass followed by the same conv kernel: without SiLU BIT FOR BIT (same arithmetic on the same fp16
    values, same summation order, incl. the zero padding of the NORMALISED tensor); with SiLU the transform uses MUFU.RCP where the
    apply pass uses a Newton reciprocal, so a few normalised values differ by one fp16 ulp; and within fp16 tolerance of the fp32
    torch reference (ResnetBlock2D: norm -> silu -> conv, /root/reference/src/utils/replace.py:239,268,321 via diffusers)."""
    Cin = C0 + C1
    a = (_rand(B, H, W, C0, seed=1) * 1.5 + 0.4).half()
    s2 = (_rand(B, H, W, C1, seed=6) * 0.7 - 0.2).half() if C1 else None
    r = _rand(B, H, W, Cout, seed=5).half() if res else None
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=2).half()
    bias = _rand(Cout, seed=3).float()
    g = _rand(Cin, seed=7).float() * 0.2 + 1.0
    bt = _rand(Cin, seed=8).float() * 0.2
    srcs = [(a, C0, C0)] + ([(s2, C1, C1)] if C1 else [])
    flat = [(t.view(B, H * W, c), c, ld) for t, c, ld in srcs]
    assert eng_mod.conv_tiles_per_image(H, W) == 2 * ((W + 7) // 8) * ((H + 31) // 32), "pick a geometry the fused kernel supports"
    slots = eng_mod.conv_tiles_per_image(H, W)
    # (1) unfused: apply pass -> normalised tensor -> conv (resident-halo swapped kernel)
    n = torch.zeros(B, H * W, Cin, dtype=torch.float16, device=DEV)
    eng_mod.k_groupnorm(flat, g, bt, n, B=B, HW=H * W, eps=1e-6, silu=silu)
    out_u = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    st_u = torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
    eng_mod.k_conv_gemm([(n.view(B, H, W, Cin), Cin, Cin)], _pack_conv_w(w), Cout, out_u, B=B, Hin=H, Win=W, ksize=3, bias=bias, out_ld=Cout,
                        out_bstride=H * W * Cout, res=(r, Cout, H * W * Cout) if res else None, stats=st_u, force_swap=2)
    # (2) fused: statistics + finalize only, the conv reads the RAW tensors
    scratch = eng_mod.k_groupnorm(flat, g, bt, None, B=B, HW=H * W, eps=1e-6, silu=silu)
    ab = eng_mod.groupnorm_ab(scratch, B, H * W, Cin)
    out_f = torch.zeros(B, H, W, Cout, dtype=torch.float16, device=DEV)
    st_f = torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
    eng_mod.k_conv_gemm(srcs, _pack_conv_w(w), Cout, out_f, B=B, Hin=H, Win=W, ksize=3, bias=bias, out_ld=Cout, out_bstride=H * W * Cout,
                        res=(r, Cout, H * W * Cout) if res else None, stats=st_f, gn_ab=ab, gn_silu=silu)
    torch.cuda.synchronize()
    if silu:
        d = (out_f.float() - out_u.float()).abs()
        assert d.max().item() <= 4e-3 and (d > 0).float().mean().item() < 0.2, (d.max().item(), (d > 0).float().mean().item())
    else:
        assert torch.equal(out_f, out_u), f"fused != unfused: {(out_f.float() - out_u.float()).abs().max().item():.3e}"
        assert torch.equal(st_f, st_u)
    # the raw inputs must not have been modified (the transform works on the shared-memory copy)
    assert torch.equal(a, (_rand(B, H, W, C0, seed=1) * 1.5 + 0.4).half())
    x = torch.cat([a, s2], -1) if C1 else a
    y = F.group_norm(x.float().permute(0, 3, 1, 2), 32, g, bt, 1e-6)
    if silu:
        y = F.silu(y)
    y = F.conv2d(y.half().float(), w.float(), bias, padding=1)
    if res:
        y = y + r.float().permute(0, 3, 1, 2)
    _close(out_f, y.permute(0, 2, 3, 1), 4e-3, 4e-3, "conv3x3 with fused GroupNorm")


# ------------------------------------------------------------------------------------------------ Upsample2D as four polyphase convs
def _poly_pack(w):
    """OIHW 3x3 -> [4 parities q = 2 py + px][O][4 taps t = 2 dy + dx][I] fp16, as Weights::conv_poly packs them: the checker-side
    restatement (oracle.sdmatte_oracle.polyphase_weights) on the fp16-rounded taps in fp32, rounded once."""
    from oracle import sdmatte_oracle as orc

    return orc.polyphase_weights(w.half().float()).half().contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [
    (1, 64, 64, 128, 128),     # single-CTA form
    (2, 128, 96, 256, 256),    # CTA-pair form (>= 74 (pixel tile, 256-channel) items), no GroupNorm in front
    (1, 64, 64, 640, 640),     # UNet level-1 upsampler: five channel tiles (no pairs), ten input slices
    (1, 160, 64, 512, 512),    # pair form, two channel-tile pairs, eight slices
])
def test_conv3x3_polyphase_upsample(eng_mod, B, H, W, Cin, Cout):
    """Upsample2D (F.interpolate(scale 2, nearest) -> conv3x3, diffusers; reached from /root/reference/src/utils/replace.py's
    up blocks) as four polyphase 2x2-tap launches over the LOW-resolution input (conv_swap_halo_kernel, p.poly): (1) against the
    same four 2x2 convs in torch with the SAME combined fp16 weights (kernel arithmetic: fp16 conv tolerance), (2) against the
    reference formulation (upsample, 3x3 conv with the original weights): the combined weights carry one extra fp16 rounding."""
    assert eng_mod.conv_can_poly(Cout, H, W)
    x = _rand(B, H, W, Cin, seed=11).half()
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=12).half()
    bias = _rand(Cout, seed=13).float()
    wq = _poly_pack(w).to(DEV)
    slots = 4 * eng_mod.conv_tiles_per_image(H, W)
    out = torch.full((B, 2 * H, 2 * W, Cout), float("nan"), dtype=torch.float16, device=DEV)
    stats = torch.full((B, slots, Cout, 2), float("nan"), device=DEV)
    for q in range(4):
        eng_mod.k_conv_gemm([(x, Cin, Cin)], wq[q], Cout, out, B=B, Hin=H, Win=W, ksize=3, bias=bias, out_ld=Cout,
                            out_bstride=4 * H * W * Cout, stats=stats, poly=q + 1)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "every output pixel is written by exactly one parity launch"
    xf = x.float().permute(0, 3, 1, 2)
    # (1) same arithmetic in torch: per parity a 2x2 correlation of the zero-padded low-resolution input
    ref1 = torch.empty(B, Cout, 2 * H, 2 * W, device=DEV)
    xp = F.pad(xf, (1, 1, 1, 1))
    for q in range(4):
        py, px = q >> 1, q & 1
        k = wq[q].float().view(Cout, 2, 2, Cin).permute(0, 3, 1, 2).contiguous()
        y = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], k, bias)
        ref1[:, :, py::2, px::2] = y
    _close(out, ref1.permute(0, 2, 3, 1), 2e-3, 2e-3, "polyphase conv vs torch with the same combined weights")
    # (2) the reference formulation
    ref2 = F.conv2d(F.interpolate(xf, scale_factor=2.0, mode="nearest"), w.float(), bias, padding=1)
    _close(out, ref2.permute(0, 2, 3, 1), 4e-3, 4e-3, "polyphase conv vs upsample -> conv3x3")
    rel = ((out.float() - ref2.permute(0, 2, 3, 1)).norm() / ref2.norm()).item()
    assert rel < 6e-4, rel  # fp16 output rounding alone is ~2.8e-4 relative RMS
    tot = stats.double().sum(1)
    o = out.double().view(B, 4 * H * W, Cout)
    assert torch.allclose(tot[..., 0], o.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(tot[..., 1], (o * o).sum(1), rtol=1e-4, atol=1e-2)


def test_conv_can_fuse_gn_is_geometry_only(eng_mod):
    assert eng_mod.conv_can_fuse_gn(3, 1, 128, 1024, 1024)
    assert eng_mod.conv_can_fuse_gn(3, 1, 128, 64, 64, has_res=1)
    assert not eng_mod.conv_can_fuse_gn(3, 2, 128, 64, 64)       # stride 2
    assert not eng_mod.conv_can_fuse_gn(1, 1, 128, 64, 64)       # 1x1
    assert not eng_mod.conv_can_fuse_gn(3, 1, 128, 64, 64, ups2=1)
    assert not eng_mod.conv_can_fuse_gn(3, 1, 320, 64, 64)       # N % 128 != 0
    assert not eng_mod.conv_can_fuse_gn(3, 1, 128, 16, 16)       # too small for 8 x 32 patches


# ------------------------------------------------------------------------------------------------ attn1 key bias
@pytest.mark.parametrize("R,B", [(64, 2), (192, 1), (512, 2)])
def test_key_bias_kernel(eng_mod, R, B):
    """key_bias_kernel against the reference chain restated with torch ops: meta_arch.py:200-204 ((tri+1)/2 -> F.interpolate(1/8,
    nearest) -> flatten), replace.py:401-403 ((1 - m) * -10000) and custom_prepare_attention_mask's nearest resize to each level's
    grid (replace.py:56-63; pinned to the reference's own function by tests/test_oracle_cpu.py::test_key_bias_matches_reference_mask_functions).
    Arbitrary trimap values (an antialiased-resized trimap is not 3-valued)."""
    import math

    g = torch.Generator(device="cpu").manual_seed(R)
    trimap = torch.rand(B, R, R, generator=g)
    trimap[:, ::16, ::16] = 1.0
    trimap[:, 8::16, ::16] = 0.5
    trimap[:, ::16, 8::16] = 0.0
    outs = eng_mod.k_key_bias(trimap.to(DEV), R)
    torch.cuda.synchronize()
    tri = trimap.unsqueeze(1) * 2 - 1
    m = F.interpolate((tri + 1) / 2, scale_factor=1 / 8, mode="nearest").flatten(start_dim=1)  # (B, S*S)
    add = ((1 - m) * -10000.0).unsqueeze(1)
    S = R // 8
    for level, got in enumerate(outs):
        s = S >> level
        want = add if level == 0 else F.interpolate(add.view(B, 1, S, S), size=(s, s), mode="nearest").view(B, 1, s * s)
        want = want.squeeze(1) * math.log2(math.e)
        assert got.shape[1] % 128 == 0 and got.shape[1] >= s * s
        torch.testing.assert_close(got[:, : s * s].cpu(), want, rtol=1e-6, atol=1e-3)
        assert torch.isinf(got[:, s * s:]).all() and (got[:, s * s:] < 0).all()
