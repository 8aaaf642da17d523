"""Upsample2D as four polyphase convs: the identity the engine's upsamplers rest on, checked on the CPU.

diffusers' Upsample2D — reached from the up blocks that /root/reference/src/utils/replace.py drives and from the VAE decoder
(meta_arch.py:255-256) — is F.interpolate(scale_factor=2, mode="nearest") followed by a 3x3 conv.  The engine never materialises
the upsampled tensor: per output parity it runs a 2x2-tap conv over the low-resolution tensor with pre-summed taps."""
import pytest
import torch
import torch.nn.functional as F


def _case(B, C, O, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(O, C, 3, 3, generator=g, dtype=torch.float64) * (9 * C) ** -0.5
    b = torch.randn(O, generator=g, dtype=torch.float64)
    return x, w, b


@pytest.mark.parametrize("B,C,O,H,W", [(1, 8, 8, 5, 7), (2, 16, 24, 8, 8), (1, 4, 4, 1, 1), (1, 3, 5, 2, 9)])
def test_polyphase_identity_is_exact(B, C, O, H, W):
    """In exact arithmetic the two formulations are the same function, borders included (the zero padding of the upsampled image
    is the zero padding of the low-resolution one): float64, agreement to rounding."""
    from oracle import sdmatte_oracle as orc

    x, w, b = _case(B, C, O, H, W, seed=H * 100 + W)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    got = orc.upsample_conv_polyphase(x, orc.polyphase_weights(w), b)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-12


def test_polyphase_fp16_weight_rounding_is_below_the_activation_rounding():
    """The engine rounds the SUM of the fp16 taps to fp16 once.  Relative RMS error of the output caused by that rounding, against the
    error caused by rounding the output itself to fp16 (which every conv of the reference's CUDA branch does): the same order and
    smaller — the polyphase form costs about one extra fp16 rounding on six of the path's ~150 layers (end to end: not visible in the
    engine-vs-fp32 error at any tap, profiles/r3f_parity.json)."""
    from oracle import sdmatte_oracle as orc

    x, w, b = _case(1, 256, 64, 16, 16, seed=3)
    w16 = w.half().double()  # the checkpoint's weights as the reference's fp16 branch uses them
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w16, b, padding=1)
    wq = orc.polyphase_weights(w16.float()).half().double()  # fp32 sums of fp16 taps, rounded to fp16 once (Weights::conv_poly)
    got = orc.upsample_conv_polyphase(x, wq, b)
    err_w = ((got - ref).norm() / ref.norm()).item()
    err_o = ((ref.half().double() - ref).norm() / ref.norm()).item()
    print(f"[polyphase] weight-rounding error {err_w:.3e}, fp16 output rounding {err_o:.3e}")
    assert err_w < err_o


def test_conv_can_poly_is_geometry_only(pkg):
    """Which upsamplers take the polyphase form is a function of (channels, low-resolution H, W) alone — never of the batch size, so a
    sample gives the same bits alone and in a batch.  At R = 1024: all three VAE decoder upsamplers and the two larger UNet ones."""
    E = pkg.engine
    for N, H, W in [(512, 128, 128), (512, 256, 256), (256, 512, 512), (1280, 32, 32), (640, 64, 64)]:
        assert E.conv_can_poly(N, H, W), (N, H, W)
    assert not E.conv_can_poly(1280, 16, 16)   # below the 8 x 32 patch of the resident-halo kernel
    assert not E.conv_can_poly(320, 64, 64)    # N % 128 != 0
    assert not E.conv_can_poly(512, 8, 8)
