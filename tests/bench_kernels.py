"""Kernel micro-benchmarks (GPU): CUDA-event timings of the hot kernels at their real shapes (B=2, 1024^2 path).
Not a test; used to compare kernel variants on the same box:  python tests/bench_kernels.py [filter]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

E = ge.load_package().engine
DEV = "cuda:0"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def attn(B, heads, Lq, Lk, bias):
    C = heads * 64
    q = torch.randn(B, Lq, C, device=DEV).half()
    k = torch.randn(B, Lk, C, device=DEV).half()
    vt = torch.randn(B, C, Lk, device=DEV).half()
    out = torch.empty(B, Lq, C, dtype=torch.float16, device=DEV)
    bz = None
    if bias:
        lv = torch.randint(0, 3, (B, Lk), device=DEV).float()
        bz = (lv * -5000.0 * math.log2(math.e)).contiguous()
    ms = timeit(lambda: E.k_attention(q, k, vt, out, B=B, heads=heads, Lq=Lq, Lk=Lk, ldq=C, ldk=C, ldvt=Lk, ldo=C, bias=bz, bias_bstride=Lk))
    fl = 4.0 * B * heads * Lq * Lk * 64
    return ms, fl / ms / 1e9


def conv(B, H, W, Cin, Cout, k=3, res=False, stats=False):
    x = torch.randn(B, H, W, Cin, device=DEV).half()
    w = (torch.randn(Cout, k * k * Cin, device=DEV) * (k * k * Cin) ** -0.5).half()
    b = torch.randn(Cout, device=DEV)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=DEV)
    r = torch.randn(B, H, W, Cout, device=DEV).half() if res else None
    st = torch.empty(B, E.conv_tiles_per_image(H, W), Cout, 2, device=DEV) if stats else None
    ms = timeit(lambda: E.k_conv_gemm([(x, Cin, Cin)], w, Cout, out, B=B, Hin=H, Win=W, ksize=k, bias=b, out_ld=Cout, out_bstride=H * W * Cout,
                                      res=(r, Cout, H * W * Cout) if res else None, stats=st))
    fl = 2.0 * B * H * W * Cout * k * k * Cin
    return ms, fl / ms / 1e9


def conv_gn(B, H, W, Cin, Cout, variant):
    """3x3 conv behind a GroupNorm+SiLU: 'apply+swap' / 'apply+swap_halo' = stand-alone apply pass + conv (both timed),
    'fused' = conv_swap_halo_kernel<true>; 'fused_nomath' / 'fused_hop' isolate the transform's shared-memory traffic / barrier hop."""
    x = torch.randn(B, H, W, Cin, device=DEV).half()
    w = (torch.randn(Cout, 9 * Cin, device=DEV) * (9 * Cin) ** -0.5).half()
    b = torch.randn(Cout, device=DEV)
    g, bt = torch.ones(Cin, device=DEV), torch.zeros(Cin, device=DEV)
    out = torch.empty(B, H, W, Cout, dtype=torch.float16, device=DEV)
    slots = E.conv_tiles_per_image(H, W)
    st = torch.empty(B, slots, Cout, 2, device=DEV)
    pre = torch.rand(B, slots, Cin, 2, device=DEV)
    pre[..., 1] += 1.0
    xf = x.view(B, H * W, Cin)
    common = dict(B=B, Hin=H, Win=W, ksize=3, bias=b, out_ld=Cout, out_bstride=H * W * Cout, stats=st)
    if variant.startswith("apply"):
        n = torch.empty_like(x)
        fs = 2 if variant.endswith("halo") else 1

        def run():
            E.k_groupnorm([(xf, Cin, Cin)], g, bt, n.view(B, H * W, Cin), B=B, HW=H * W, eps=1e-6, silu=1, pre=(pre, None), pre_slots=slots)
            E.k_conv_gemm([(n, Cin, Cin)], w, Cout, out, force_swap=fs, **common)
    else:
        mode = {"fused": 1, "fused_hop": 2, "fused_nomath": 4, "fused_hop_noepilogue": 10, "fused_hop_nostores": 18,
                "bisect_hotx": 2 + 32, "bisect_notma": 2 + 128, "bisect_notma_noepilogue": 2 + 128 + 8, "bisect_hotx_noepilogue": 2 + 32 + 8}[variant]
        scratch = E.k_groupnorm([(xf, Cin, Cin)], g, bt, None, B=B, HW=H * W, eps=1e-6, silu=1, pre=(pre, None), pre_slots=slots)
        ab = E.groupnorm_ab(scratch, B, H * W, Cin)

        def run():
            E.k_groupnorm([(xf, Cin, Cin)], g, bt, None, B=B, HW=H * W, eps=1e-6, silu=1, pre=(pre, None), pre_slots=slots)
            E.k_conv_gemm([(x, Cin, Cin)], w, Cout, out, gn_ab=ab, gn_silu=mode, **common)
    ms = timeit(run)
    return ms, 2.0 * B * H * W * Cout * 9 * Cin / ms / 1e9


def linear(B, M, N, K, res=False):
    x = torch.randn(B, M, K, device=DEV).half()
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).half()
    b = torch.randn(N, device=DEV)
    out = torch.randn(B, M, N, device=DEV).half()
    ms = timeit(lambda: E.k_conv_gemm([(x, K, K)], w, N, out, B=B, Hin=1, Win=M, bias=b, out_ld=N, out_bstride=M * N,
                                      res=(out, N, M * N) if res else None))
    return ms, 2.0 * B * M * N * K / ms / 1e9


def geglu(B, M, C):
    """GEGLU projection of a transformer block's feed-forward (EPI_GEGLU): [M, C] x [8C, C]^T -> [M, 4C]."""
    x = torch.randn(B, M, C, device=DEV).half()
    w = (torch.randn(8 * C, C, device=DEV) * C ** -0.5).half()
    b = torch.randn(8 * C, device=DEV)
    out = torch.empty(B, M, 4 * C, device=DEV).half()
    ms = timeit(lambda: E.k_conv_gemm([(x, C, C)], w, 8 * C, out, B=B, Hin=1, Win=M, mode=2, bias=b, out_ld=4 * C, out_bstride=M * 4 * C))
    return ms, 2.0 * B * M * 8 * C * C / ms / 1e9


def linear_t(B, M, N, K):
    """to_v with the transposed store (EPI_F16_T): out[b][n][m]."""
    x = torch.randn(B, M, K, device=DEV).half()
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).half()
    out = torch.empty(B, N, M, device=DEV).half()
    ms = timeit(lambda: E.k_conv_gemm([(x, K, K)], w, N, out, B=B, Hin=1, Win=M, mode=1, out_ld=M, out_bstride=N * M))
    return ms, 2.0 * B * M * N * K / ms / 1e9


def gn(B, H, W, C):
    """GroupNorm(+SiLU) with the statistics arriving as conv-epilogue partials (the in-engine case): reduce + finalize + apply."""
    x = torch.randn(B, H * W, C, device=DEV).half()
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    out = torch.empty_like(x)
    slots = E.conv_tiles_per_image(H, W)
    pre = torch.rand(B, slots, C, 2, device=DEV)
    pre[..., 1] += 1.0
    ms = timeit(lambda: E.k_groupnorm([(x, C, C)], g, b, out, B=B, HW=H * W, eps=1e-6, silu=1, pre=(pre, None), pre_slots=slots))
    return ms, 4.0 * B * H * W * C / ms / 1e9  # "TFLOP/s" column = TB/s here


def scores(B, L, D):
    """VAE mid-block attention scores: fp32 [B][L][L] = scale * Q K^T (batched weights, EPI_F32)."""
    q = torch.randn(B, L, D, device=DEV).half()
    k = torch.randn(B, L, D, device=DEV).half()
    out = torch.empty(B, L, L, dtype=torch.float32, device=DEV)
    ms = timeit(lambda: E.k_conv_gemm([(q, D, D)], k, L, out, B=B, Hin=1, Win=L, mode=3, scale=D ** -0.5, w_bstride=L * D, out_ld=L,
                                      out_bstride=L * L), iters=5, warm=2)
    return ms, 2.0 * B * L * L * D / ms / 1e9


CASES = {
    "vae_qk scores 16384x16384x512 B2 f32": lambda: scores(2, 16384, 512),
    "gn+silu 128ch @1024^2 B4 (TB/s)": lambda: gn(4, 1024, 1024, 128),
    "gn+silu 256ch @512^2 B4 (TB/s)": lambda: gn(4, 512, 512, 256),
    "gn+silu 320ch @128^2 B8 (TB/s)": lambda: gn(8, 128, 128, 320),
    "attn_self_L0 (B2 h5 16384x16384 bias)": lambda: attn(2, 5, 16384, 16384, True),
    "attn_cross_L0 (B2 h5 16384x16384)": lambda: attn(2, 5, 16384, 16384, False),
    "attn_cross_L1 (B2 h10 4096x16384)": lambda: attn(2, 10, 4096, 16384, False),
    "attn_self_L2 (B2 h20 1024x1024 bias)": lambda: attn(2, 20, 1024, 1024, True),
    **{f"gn+conv3x3 128->128 @1024^2 B4 {v}": (lambda v=v: conv_gn(4, 1024, 1024, 128, 128, v)) for v in ("apply+swap", "apply+swap_halo", "fused", "fused_nomath", "fused_hop", "fused_hop_noepilogue", "fused_hop_nostores")},
    **{f"gn+conv3x3 256->256 @512^2 B4 {v}": (lambda v=v: conv_gn(4, 512, 512, 256, 256, v)) for v in ("apply+swap_halo", "fused", "fused_nomath", "fused_hop")},
    **{f"gn+conv3x3 512->512 @256^2 B4 {v}": (lambda v=v: conv_gn(4, 256, 256, 512, 512, v)) for v in ("apply+swap_halo", "fused", "fused_nomath", "fused_hop")},
    **{f"bisect 256->256 @512^2 B4 {v}": (lambda v=v: conv_gn(4, 512, 512, 256, 256, v)) for v in (
        "fused_hop", "bisect_hotx", "bisect_notma", "fused_hop_noepilogue", "bisect_hotx_noepilogue", "bisect_notma_noepilogue")},
    **{f"bisect 128->128 @1024^2 B4 {v}": (lambda v=v: conv_gn(4, 1024, 1024, 128, 128, v)) for v in (
        "fused_hop", "bisect_hotx", "bisect_notma", "fused_hop_noepilogue", "bisect_notma_noepilogue")},
    "conv3x3 128->128 @1024^2 B2": lambda: conv(2, 1024, 1024, 128, 128),
    "conv3x3 128->128 @1024^2 B2 +res+stats": lambda: conv(2, 1024, 1024, 128, 128, res=True, stats=True),
    "conv3x3 256->256 @512^2 B4": lambda: conv(4, 512, 512, 256, 256),
    "conv3x3 256->256 @512^2 B4 +res+stats": lambda: conv(4, 512, 512, 256, 256, res=True, stats=True),
    "conv3x3 512->512 @256^2 B4": lambda: conv(4, 256, 256, 512, 512),
    "conv3x3 320->320 @128^2 B8 +stats": lambda: conv(8, 128, 128, 320, 320, stats=True),
    "conv1x1 128->256 @512^2 B4": lambda: conv(4, 512, 512, 128, 256, k=1),
    "conv1x1 64->128 @1024^2 B4 +stats (im2col conv_in)": lambda: conv(4, 1024, 1024, 64, 128, k=1, stats=True),
    "conv1x1 64->128 @1024^2 B4": lambda: conv(4, 1024, 1024, 64, 128, k=1),
    "conv1x1 256->128 @1024^2 B2 (vae shortcut)": lambda: conv(2, 1024, 1024, 256, 128, k=1),
    "geglu 16384x320 B8": lambda: geglu(8, 16384, 320),
    "geglu 4096x640 B8": lambda: geglu(8, 4096, 640),
    "linear_T 16384x320x1024 B8 (cross V^T)": lambda: linear_t(8, 16384, 320, 1024),
    "linear_T 16384x320x320 B8 (self V^T)": lambda: linear_t(8, 16384, 320, 320),
    "linear 16384x320x320 B8 +res": lambda: linear(8, 16384, 320, 320, res=True),
    "linear 16384x320x1024 B8 (cross K)": lambda: linear(8, 16384, 320, 1024),
    "linear 4096x640x640 B8": lambda: linear(8, 4096, 640, 640),
    "linear 16384x320x1280 B8 +res (ff out)": lambda: linear(8, 16384, 320, 1280, res=True),
}

if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    print(f"{'case':48s} {'ms':>9s} {'TFLOP/s':>9s}")
    for name, fn in CASES.items():
        if flt and flt not in name:
            continue
        ms, tf = fn()
        print(f"{name:48s} {ms:9.3f} {tf:9.1f}")
