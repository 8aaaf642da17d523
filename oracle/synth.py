"""TEST INFRASTRUCTURE: synthetic checkpoint (reference key names) and synthetic inputs.

No real SDMatte*.safetensors is available offline, so parity is measured with a seeded, variance-preserving random
checkpoint of the exact architecture (random N(0,1) weights through 60+ layers would saturate alpha to 0/1 and make the
comparison vacuous).  Every tensor is generated from its own seed (base seed + crc32(name)) so the checkpoint is
identical on every machine regardless of enumeration order.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

UNET_CH = (320, 640, 1280, 1280)


def param_shapes(include_unused: bool = True) -> "OrderedDict[str, Tuple[int, ...]]":
    """Every parameter of CustomUNet (replace.py:184-362 + utils.py:13-41) and the SD VAE, with the reference's key names."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(n, o, i, k):
        s[n + ".weight"] = (o, i, k, k)
        s[n + ".bias"] = (o,)

    def lin(n, o, i, bias=True):
        s[n + ".weight"] = (o, i)
        if bias:
            s[n + ".bias"] = (o,)

    def norm(n, c):
        s[n + ".weight"] = (c,)
        s[n + ".bias"] = (c,)

    def resnet(p, cin, cout, temb):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        if temb:
            lin(p + ".time_emb_proj", cout, 1280)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def transformer(p, c):
        norm(p + ".norm", c)
        lin(p + ".proj_in", c, c)
        t = p + ".transformer_blocks.0"
        for a, kv in (("attn1", c), ("attn2", 1024)):
            norm(f"{t}.norm{1 if a == 'attn1' else 2}", c)
            lin(f"{t}.{a}.to_q", c, c, False)
            lin(f"{t}.{a}.to_k", c, kv, False)
            lin(f"{t}.{a}.to_v", c, kv, False)
            lin(f"{t}.{a}.to_out.0", c, c)
        norm(t + ".norm3", c)
        lin(t + ".ff.net.0.proj", 8 * c, c)
        lin(t + ".ff.net.2", c, 4 * c)
        lin(p + ".proj_out", c, c)

    # ---- UNet
    conv("unet.conv_in", 320, 8, 3)  # 8 input channels after replace_unet_conv_in(unet, 2), utils.py:13-30
    conv("unet.aux_conv_in", 1024, 4, 3)  # utils.py:33-41
    for n, i in (("time_embedding", 320), ("bbox_embedding", 1280)) + ((("point_embedding", 1680),) if include_unused else ()):
        lin(f"unet.{n}.linear_1", 1280, i)
        lin(f"unet.{n}.linear_2", 1280, 1280)
    cin = 320
    for i, c in enumerate(UNET_CH):
        for j in range(2):
            resnet(f"unet.down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, True)
            if i < 3:
                transformer(f"unet.down_blocks.{i}.attentions.{j}", c)
        cin = c
        if i < 3:
            conv(f"unet.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
    resnet("unet.mid_block.resnets.0", 1280, 1280, True)
    transformer("unet.mid_block.attentions.0", 1280)
    resnet("unet.mid_block.resnets.1", 1280, 1280, True)
    skips = [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
    prev = 1280
    for i, c in enumerate((1280, 1280, 640, 320)):
        for j in range(3):
            resnet(f"unet.up_blocks.{i}.resnets.{j}", prev + skips.pop(), c, True)
            prev = c
            if i > 0:
                transformer(f"unet.up_blocks.{i}.attentions.{j}", c)
        if i < 3:
            conv(f"unet.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("unet.conv_norm_out", 320)
    conv("unet.conv_out", 4, 320, 3)

    # ---- VAE
    def vae_attn(p):
        norm(p + ".group_norm", 512)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{p}.{n}", 512, 512)

    e = "vae.encoder"
    conv(e + ".conv_in", 128, 3, 3)
    cin = 128
    for i, c in enumerate((128, 256, 512, 512)):
        for j in range(2):
            resnet(f"{e}.down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, False)
        cin = c
        if i < 3:
            conv(f"{e}.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
    resnet(e + ".mid_block.resnets.0", 512, 512, False)
    vae_attn(e + ".mid_block.attentions.0")
    resnet(e + ".mid_block.resnets.1", 512, 512, False)
    norm(e + ".conv_norm_out", 512)
    conv(e + ".conv_out", 8, 512, 3)
    conv("vae.quant_conv", 8, 8, 1)
    conv("vae.post_quant_conv", 4, 4, 1)
    d = "vae.decoder"
    conv(d + ".conv_in", 512, 4, 3)
    resnet(d + ".mid_block.resnets.0", 512, 512, False)
    vae_attn(d + ".mid_block.attentions.0")
    resnet(d + ".mid_block.resnets.1", 512, 512, False)
    cin = 512
    for i, c in enumerate((512, 512, 256, 128)):
        for j in range(3):
            resnet(f"{d}.up_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, False)
        cin = c
        if i < 3:
            conv(f"{d}.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm(d + ".conv_norm_out", 128)
    conv(d + ".conv_out", 3, 128, 3)
    return s


def _gain(name: str) -> float:
    damp = (".conv2.weight", ".to_out.0.weight", ".ff.net.2.weight", ".proj_out.weight")
    return 0.5 if name.endswith(damp) else 1.0


def make_checkpoint(seed: int = 1234, dtype=torch.float32, include_unused: bool = False) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(include_unused).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2**63))
        if len(shape) == 1:
            is_norm_w = name.endswith(".weight")
            t = torch.randn(shape, generator=g) * 0.05
            if is_norm_w:
                t = t + 1.0
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * (_gain(name) / fan_in ** 0.5)
        sd[name] = t.to(dtype)
    return sd


def make_inputs(B: int, R: int, seed: int = 0, Hin: int | None = None, Win: int | None = None):
    """image (B,H,W,3) fp32 in [0,1] (white noise blended 50/50 with a smooth field) and a trimap (B,H,W) in {0, .5, 1}
    (~30% fg / 20% unknown / 50% bg, guaranteed to contain all three values)."""
    H, W = Hin or R, Win or R
    g = torch.Generator().manual_seed(seed)
    noise = torch.rand(B, H, W, 3, generator=g)
    low = torch.rand(B, 3, max(2, H // 64 + 1), max(2, W // 64 + 1), generator=g)
    smooth = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    image = (0.5 * noise + 0.5 * smooth).clamp(0, 1).contiguous()
    blob = torch.rand(B, 1, max(2, H // 48 + 2), max(2, W // 48 + 2), generator=g)
    field = F.interpolate(blob, size=(H, W), mode="bicubic", align_corners=True).squeeze(1)
    trimap = torch.zeros(B, H, W)
    for b in range(B):
        q50, q70 = torch.quantile(field[b].flatten()[:: max(1, (H * W) // 65536)], torch.tensor([0.5, 0.7]))
        trimap[b] = torch.where(field[b] > q70, 1.0, torch.where(field[b] > q50, 0.5, 0.0))
        cy, cx = H // 2, W // 2
        trimap[b, cy - 8: cy + 8, cx - 8: cx + 8] = 1.0  # at least one foreground key at every UNet level
        trimap[b, :8, :8] = 0.0
        trimap[b, :8, 8:16] = 0.5
    return image, trimap.contiguous()
