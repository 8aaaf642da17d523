"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU/torch restatement of the reference SDMatte matte path.

Nothing on the product path may import this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do (as the checker / the timed CPU baseline).

PARITY STATUS: *partially pinned*.  The reference (flybirdxx/ComfyUI-SDMatte @ /root/reference) ships no tests, no
golden vectors and no checkpoints, and its arithmetic lives in the un-vendored dependency `diffusers>=0.25.0`
(requirements.txt:1; not installed here, no network).  What IS pinned against the reference's own code, executed in
this container from /root/reference (see oracle/ref_lifted.py and tests/golden/make_golden.py):
  * custom_prepare_attention_mask / custom_get_attention_scores  (src/utils/replace.py:20-122)
  * replace_unet_conv_in / add_aux_conv_in                       (src/utils/utils.py:13-41)
  * _resize_norm_image_bchw / _resize_mask_b1hw, mask_refine + output composition (sdmatte_nodes.py:204-214,365-397)
What is restated from the published diffusers algorithm and therefore "parity unpinned":
  ResnetBlock2D, Transformer2DModel/BasicTransformerBlock/Attention/GEGLU, Down/Upsample2D, AutoencoderKL
  encoder/decoder, get_timestep_embedding/TimestepEmbedding.  Structural self-check available without diffusers:
  the enumerated parameter shapes reproduce the published counts (SD-2.1 UNet 865.91 M, SD VAE 83.65 M), see
  tests/test_oracle_cpu.py::test_parameter_counts.

The functions below follow, line by line:
  SDMatte.forward        /root/reference/src/modeling/SDMatte/meta_arch.py:127-261
  CustomUNet.forward     /root/reference/src/utils/replace.py:379-549 (module tree :184-362)
with node flags frozen as sdmatte_nodes.py:286-296 sets them (aux_input="trimap", use_coor_input=True, ...).
State-dict keys are the reference's (`unet.*`, `vae.*`), so a real SDMatte.safetensors drops in.

mode="fp32"    : everything in fp32 (what the reference's CPU branch computes, sdmatte_nodes.py:359-360)
mode="fp16sim" : fp32 math with the fp16 rounding points of the reference's CUDA autocast path emulated
mode="autocast": the reference's real CUDA branch (sdmatte_nodes.py:355-358): the same graph on a CUDA device under a
                 genuine `torch.autocast("cuda", dtype=torch.float16)` with fp32 master weights (cuDNN / cuBLAS kernels — allowed
                 for the checker, never for the product), UNet attention = SlicedAttnProcessor(slice_size=1) semantics
                 (sdmatte_nodes.py:331-335): one (sample, head) slice at a time through `torch.baddbmm` + softmax + `torch.bmm`
                 exactly as custom_get_attention_scores (replace.py:75-122); VAE attention = SDPA (diffusers default processor).
`device`: where the graph runs ("cpu" default; "cuda" for the GPU-resident checker and the GPU baseline of bench.py).  fp32 on a
CUDA device runs with TF32 disabled.  `capture`: receives, besides the outputs, one tensor per block ("taps", graph order kept in
capture["_order"]) for the per-block error-growth curves of tests/test_parity_gpu.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SCALING_FACTOR = 0.18215
UNET_CH = (320, 640, 1280, 1280)
UNET_HEADS = (5, 10, 20, 20)
VAE_CH = (128, 256, 512, 512)


class _Ctx:
    def __init__(self, sd, mode, sliced, device="cpu", capture=None):
        self.sd = sd
        self.mode = mode
        self.sliced = sliced
        self.device = torch.device(device)
        self.capture = capture

    def tap(self, name, x):
        """Record a block output (NCHW) for the per-block parity curves; no-op unless a capture dict was given."""
        if self.capture is not None:
            self.capture[name] = x.detach()
            self.capture.setdefault("_order", []).append(name)

    def r(self, x):  # fp16 rounding point of the autocast path
        return x.half().float() if self.mode == "fp16sim" else x

    def w(self, name):
        t = self.sd[name]
        t = t.float()
        return t.half().float() if self.mode == "fp16sim" else t


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0) as called at meta_arch.py:181-186."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def _linear(c, x, name, bias=True):
    return c.r(F.linear(c.r(x), c.w(name + ".weight"), c.sd[name + ".bias"].float() if bias else None))


def _conv(c, x, name, stride=1, padding=1):
    return c.r(F.conv2d(c.r(x), c.w(name + ".weight"), c.sd[name + ".bias"].float(), stride=stride, padding=padding))


def _gn(c, x, name, eps):
    return F.group_norm(x, 32, c.sd[name + ".weight"].float(), c.sd[name + ".bias"].float(), eps)


def _ln(c, x, name):
    return F.layer_norm(x, (x.shape[-1],), c.sd[name + ".weight"].float(), c.sd[name + ".bias"].float(), 1e-5)


def _time_mlp(c, x, name):  # TimestepEmbedding: linear_2(silu(linear_1(x)))
    return _linear(c, F.silu(_linear(c, x, name + ".linear_1")), name + ".linear_2")


def _resnet(c, x, p, emb, eps):
    """diffusers ResnetBlock2D."""
    h = _conv(c, F.silu(_gn(c, x, p + ".norm1", eps)), p + ".conv1")
    if emb is not None:
        t = _linear(c, F.silu(emb), p + ".time_emb_proj")
        h = c.r(h + t[:, :, None, None])
    h = _conv(c, F.silu(_gn(c, h, p + ".norm2", eps)), p + ".conv2")
    if (p + ".conv_shortcut.weight") in c.sd:
        x = _conv(c, x, p + ".conv_shortcut", padding=0)
    return c.r(x + h)


def prepare_key_bias(attention_mask: torch.Tensor, target_length: int) -> torch.Tensor:
    """custom_prepare_attention_mask (replace.py:20-72) for out_dim=3: nearest resize of the (B,1,L0) additive bias."""
    current = attention_mask.shape[-1]
    if current != target_length:
        B = attention_mask.shape[0]
        cs, ts = int(math.sqrt(current)), int(math.sqrt(target_length))
        assert cs * cs == current and ts * ts == target_length
        attention_mask = F.interpolate(attention_mask.view(B, -1, cs, cs), size=(ts, ts), mode="nearest").view(B, 1, target_length)
    return attention_mask


def _attention(c, x, ctx, p, heads, key_bias):
    """diffusers Attention with AttnProcessor / SlicedAttnProcessor(1) + custom_get_attention_scores (replace.py:75-122)."""
    B, L, C = x.shape
    q = _linear(c, x, p + ".to_q", bias=False)
    k = _linear(c, ctx, p + ".to_k", bias=False)
    v = _linear(c, ctx, p + ".to_v", bias=False)
    Lk = k.shape[1]
    d = C // heads
    q = q.view(B, L, heads, d).transpose(1, 2)
    k = k.view(B, Lk, heads, d).transpose(1, 2)
    v = v.view(B, Lk, heads, d).transpose(1, 2)
    scale = d ** -0.5
    bias = None
    if key_bias is not None:
        bias = prepare_key_bias(key_bias, Lk)[:, :, None, :]  # (B,1,1,Lk) -> every head, every query row
    if c.mode == "autocast":
        # SlicedAttnProcessor(slice_size=1): the (B*heads) slices one at a time; per slice replace.py:92-120 verbatim in effect:
        # baddbmm(mask or empty, q, k^T, beta, alpha=scale) [autocast: fp16 out] -> softmax [autocast: fp32] -> .to(q.dtype) -> bmm
        out = torch.zeros(B, heads, L, d, device=q.device, dtype=q.dtype)
        for b in range(B):
            for h in range(heads):
                qs, ks, vs = q[b, h:h + 1], k[b, h:h + 1], v[b, h:h + 1]
                if bias is not None:
                    s = torch.baddbmm(bias[b], qs, ks.transpose(-1, -2), beta=1, alpha=scale)
                else:
                    empty = torch.empty(1, L, Lk, dtype=qs.dtype, device=qs.device)
                    s = torch.baddbmm(empty, qs, ks.transpose(-1, -2), beta=0, alpha=scale)
                pr = s.softmax(dim=-1).to(qs.dtype)
                del s
                out[b, h:h + 1] = torch.bmm(pr, vs)
                del pr
        out = out.transpose(1, 2).reshape(B, L, C)
        return _linear(c, out, p + ".to_out.0")
    out = torch.empty(B, heads, L, d, device=q.device)
    for b in range(B):
        for h in range(heads if c.sliced else 1):
            hs = slice(h, h + 1) if c.sliced else slice(None)
            s = torch.matmul(q[b, hs], k[b, hs].transpose(-1, -2)) * scale  # baddbmm(beta, alpha=scale)
            if bias is not None:
                s = s + c.r(bias[b])  # autocast casts the additive mask to fp16 too (exact for 0 / -5000 / -10000)
            s = c.r(s)
            pr = c.r(s.softmax(dim=-1))
            out[b, hs] = c.r(torch.matmul(pr, v[b, hs]))
    out = out.transpose(1, 2).reshape(B, L, C)
    return _linear(c, out, p + ".to_out.0")


def _transformer(c, x, p, heads, ctx, key_bias):
    """Transformer2DModel(use_linear_projection=True) with one BasicTransformerBlock."""
    B, C, H, W = x.shape
    res = x
    h = _gn(c, x, p + ".norm", 1e-6).permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _linear(c, h, p + ".proj_in")
    t = p + ".transformer_blocks.0"
    h = c.r(_attention(c, _ln(c, h, t + ".norm1"), _ln(c, h, t + ".norm1"), t + ".attn1", heads, key_bias) + h)
    h = c.r(_attention(c, _ln(c, h, t + ".norm2"), ctx, t + ".attn2", heads, None) + h)
    g = _linear(c, _ln(c, h, t + ".norm3"), t + ".ff.net.0.proj")
    a, gate = g.chunk(2, dim=-1)
    ff = c.r(a * c.r(F.gelu(gate)))
    h = c.r(_linear(c, ff, t + ".ff.net.2") + h)
    h = _linear(c, h, p + ".proj_out").reshape(B, H, W, C).permute(0, 3, 1, 2)
    return c.r(h + res)


def unet_forward(c, sample, trans, ctx, coords_emb, attention_mask, capture=None, point_prompt=False):
    """CustomUNet.forward (replace.py:379-549) with timestep=None.  `point_prompt`: added_cond_kwargs carries "point_coords"
    (-> point_embedding, replace.py:446-450) instead of "bbox_mask_coords" (-> bbox_embedding, :451-455)."""
    B = sample.shape[0]
    key_bias = ((1 - attention_mask) * -10000.0).unsqueeze(1)  # replace.py:401-403
    op_emb = _time_mlp(c, timestep_embedding(trans.float(), 320), "unet.time_embedding")  # :430-435
    aug_emb = _time_mlp(c, coords_emb.reshape(B, -1), "unet.point_embedding" if point_prompt else "unet.bbox_embedding")
    emb = c.r(op_emb + aug_emb)  # :459
    x = _conv(c, sample, "unet.conv_in")  # :462
    c.tap("unet.conv_in", x)
    skips = [x]
    for i in range(4):
        bp = f"unet.down_blocks.{i}"
        for j in range(2):
            x = _resnet(c, x, f"{bp}.resnets.{j}", emb, 1e-5)
            if i < 3:
                x = _transformer(c, x, f"{bp}.attentions.{j}", UNET_HEADS[i], ctx, key_bias)
            c.tap(f"unet.down{i}.{j}", x)
            skips.append(x)
        if i < 3:
            x = _conv(c, x, f"{bp}.downsamplers.0.conv", stride=2, padding=1)
            c.tap(f"unet.down{i}.ds", x)
            skips.append(x)
    x = _resnet(c, x, "unet.mid_block.resnets.0", emb, 1e-5)
    x = _transformer(c, x, "unet.mid_block.attentions.0", 20, ctx, key_bias)
    x = _resnet(c, x, "unet.mid_block.resnets.1", emb, 1e-5)
    c.tap("unet.mid", x)
    rheads = (20, 20, 10, 5)
    for i in range(4):
        bp = f"unet.up_blocks.{i}"
        for j in range(3):
            x = torch.cat([x, skips.pop()], dim=1)
            x = _resnet(c, x, f"{bp}.resnets.{j}", emb, 1e-5)
            if i > 0:
                x = _transformer(c, x, f"{bp}.attentions.{j}", rheads[i], ctx, key_bias)
            if not (i < 3 and j == 2):
                c.tap(f"unet.up{i}.{j}", x)
        if i < 3:
            c.tap(f"unet.up{i}.2.lo", x)  # before the nearest x2 (the engine's polyphase upsampler keeps the tensor at low resolution)
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            c.tap(f"unet.up{i}.2", x)  # AFTER the nearest x2 (elsewhere the engine fuses the upsampling into the block's store)
            x = _conv(c, x, f"{bp}.upsamplers.0.conv")
            c.tap(f"unet.up{i}.us", x)
    assert not skips
    x = F.silu(_gn(c, x, "unet.conv_norm_out", 1e-5))
    return _conv(c, x, "unet.conv_out")


def _vae_attention(c, x, p):
    """diffusers Attention(heads=1, residual_connection=True, norm_num_groups=32) in the VAE mid block."""
    B, C, H, W = x.shape
    h = _gn(c, x, p + ".group_norm", 1e-6).view(B, C, H * W).transpose(1, 2)
    q, k, v = _linear(c, h, p + ".to_q"), _linear(c, h, p + ".to_k"), _linear(c, h, p + ".to_v")
    out = torch.empty_like(q)
    for b in range(B):
        if c.mode == "autocast":  # AttnProcessor2_0: F.scaled_dot_product_attention on the fp16 q/k/v, one head of 512
            out[b] = F.scaled_dot_product_attention(q[b][None, None], k[b][None, None], v[b][None, None])[0, 0]
            continue
        s = torch.matmul(q[b], k[b].transpose(0, 1)) * (C ** -0.5)
        out[b] = c.r(torch.matmul(s.softmax(dim=-1), v[b]))
        del s
    out = _linear(c, out, p + ".to_out.0").transpose(1, 2).reshape(B, C, H, W)
    return c.r(out + x)


def vae_encode(c, x, tag="enc"):
    """AutoencoderKL.encoder + quant_conv, mean half, * scaling_factor (meta_arch.py:142-145,209-212)."""
    e = "vae.encoder"
    h = _conv(c, x, e + ".conv_in")
    c.tap(f"{tag}.conv_in", h)
    for i in range(4):
        for j in range(2):
            h = _resnet(c, h, f"{e}.down_blocks.{i}.resnets.{j}", None, 1e-6)
        if i < 3:
            h = _conv(c, F.pad(h, (0, 1, 0, 1)), f"{e}.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0)
        c.tap(f"{tag}.down{i}", h)
    h = _resnet(c, h, e + ".mid_block.resnets.0", None, 1e-6)
    h = _vae_attention(c, h, e + ".mid_block.attentions.0")
    c.tap(f"{tag}.mid_attn", h)
    h = _resnet(c, h, e + ".mid_block.resnets.1", None, 1e-6)
    c.tap(f"{tag}.mid", h)
    h = _conv(c, F.silu(_gn(c, h, e + ".conv_norm_out", 1e-6)), e + ".conv_out")
    moments = _conv(c, h, "vae.quant_conv", padding=0)
    mean, _ = torch.chunk(moments, 2, dim=1)
    return c.r(mean * SCALING_FACTOR)


def vae_decode(c, z):
    """post_quant_conv + AutoencoderKL.decoder (meta_arch.py:255-256)."""
    d = "vae.decoder"
    h = _conv(c, z, "vae.post_quant_conv", padding=0)
    h = _conv(c, h, d + ".conv_in")
    c.tap("dec.conv_in", h)
    h = _resnet(c, h, d + ".mid_block.resnets.0", None, 1e-6)
    h = _vae_attention(c, h, d + ".mid_block.attentions.0")
    h = _resnet(c, h, d + ".mid_block.resnets.1", None, 1e-6)
    c.tap("dec.mid", h)
    for i in range(4):
        for j in range(3):
            h = _resnet(c, h, f"{d}.up_blocks.{i}.resnets.{j}", None, 1e-6)
        if i < 3:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(c, h, f"{d}.up_blocks.{i}.upsamplers.0.conv")
        c.tap(f"dec.up{i}", h)
    return _conv(c, F.silu(_gn(c, h, d + ".conv_norm_out", 1e-6)), d + ".conv_out")


def polyphase_weights(w: torch.Tensor) -> torch.Tensor:
    """Checker-side restatement of the engine's upsampler weight packing (csrc/engine.cu Weights::conv_poly): a 3x3 conv over a
    nearest-x2 upsampled image (diffusers Upsample2D: F.interpolate(scale_factor=2, mode="nearest") then conv) equals, for each
    output parity (py, px), a 2x2-tap conv over the LOW-resolution image whose taps are the sums of the 3x3 taps that land on the
    same input pixel: rows {0 | 1,2} for py = 0, {0,1 | 2} for py = 1 (columns alike).  OIHW -> [4 (q = 2 py + px)][O][4 (t = 2 dy + dx)][I],
    in the dtype of `w` (the engine sums the fp16-rounded taps in fp32 and rounds the sum to fp16 once)."""
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = []
    for q in range(4):
        py, px = q >> 1, q & 1
        taps = []
        for t in range(4):
            dy, dx = t >> 1, t & 1
            acc = 0
            for ky in rows[py][dy]:
                for kx in rows[px][dx]:
                    acc = acc + w[:, :, ky, kx]
            taps.append(acc)
        out.append(torch.stack(taps, 1))
    return torch.stack(out)


def upsample_conv_polyphase(x: torch.Tensor, wq: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """The four polyphase convs with the weights of polyphase_weights(): x [B][C][H][W] -> [B][O][2H][2W]."""
    B, C, H, W = x.shape
    O = wq.shape[1]
    xp = F.pad(x, (1, 1, 1, 1))
    out = x.new_empty(B, O, 2 * H, 2 * W)
    for q in range(4):
        py, px = q >> 1, q & 1
        k = wq[q].reshape(O, 2, 2, C).permute(0, 3, 1, 2).contiguous()
        out[:, :, py::2, px::2] = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], k, bias)
    return out


def to_device(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """fp32 master copy of the checkpoint on `device` (what `.to(device)` leaves at sdmatte_nodes.py:323)."""
    return {k: v.float().to(device) for k, v in sd.items() if k.startswith(("unet.", "vae."))}


@torch.no_grad()
def point_coords_embedding(coor: torch.Tensor) -> torch.Tensor:
    """meta_arch.py:153-176: N point coordinates per sample are zero-padded to the first i >= N that divides 1680, each embedded
    with 1680 / i sinusoid channels -> (B, 1680), the input width of point_embedding (meta_arch.py:107-108)."""
    B, N = coor.shape
    for i in range(N, 1680):
        if 1680 % i == 0:
            num_channels = 1680 // i
            coor = torch.cat([coor, torch.zeros((B, i - N), dtype=coor.dtype, device=coor.device)], dim=1)
            break
    else:
        raise ValueError("too many point coordinates")
    half = num_channels // 2
    # diffusers get_timestep_embedding(dim, flip_sin_to_cos=True, downscale_freq_shift=0); an odd dim is zero-padded by one column
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=coor.device) / half
    emb = coor.flatten()[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)
    if num_channels % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb.reshape(B, -1)


def forward(sd: Dict[str, torch.Tensor], image: torch.Tensor, trimap: torch.Tensor, is_transparent=False,
            mode: str = "fp32", sliced: bool = False, capture: Optional[dict] = None, device="cpu",
            prompt: str = "trimap", coords: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """image (B,R,R,3) fp32 in [0,1] and trimap (B,R,R) fp32 in [0,1], both ALREADY at inference size R.

    `prompt` (meta_arch.py:22-28,130-197): which visual prompt the auxiliary image is — "trimap" (the node's
    case, coords fixed to [0,0,1,1] by sdmatte_nodes.py:353), "mask" / "bbox_mask" (4 coordinates per sample through
    bbox_embedding) or "point_mask" (N point coordinates per sample through point_embedding); `coords` (B, 4) / (B, N).  The
    auxiliary image takes the trimap's place everywhere (VAE latent, cross-attention context, attention-mask source).

    Returns {"alpha": (B,1,R,R) in [0,1], "label_mean": pre-clip decoder channel mean, + intermediates} on `device`.
    Follows sdmatte_nodes.py:339-360 (pre-processing at native size) then meta_arch.py:127-261.
    `sd` must already live on `device` (see to_device) — the reference moves the model once, before the forward.
    """
    device = torch.device(device)
    if mode == "autocast":
        assert device.type == "cuda", "mode='autocast' is the reference's CUDA branch (sdmatte_nodes.py:355-358)"
        with torch.autocast(device_type="cuda", dtype=torch.float16):
            return _forward(sd, image, trimap, is_transparent, mode, True, capture, device, prompt, coords)
    if device.type == "cuda":  # fp32 checker on the GPU: no TF32 anywhere
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        try:
            return _forward(sd, image, trimap, is_transparent, mode, sliced, capture, device, prompt, coords)
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return _forward(sd, image, trimap, is_transparent, mode, sliced, capture, device, prompt, coords)


def _forward(sd, image, trimap, is_transparent, mode, sliced, capture, device, prompt="trimap", coords=None):
    c = _Ctx(sd, mode, sliced, device, capture)
    B = image.shape[0]
    image, trimap = image.to(device), trimap.to(device)
    rgb = (image.permute(0, 3, 1, 2).float() - 0.5) / 0.5  # Normalize(0.5, 0.5), sdmatte_nodes.py:204-209
    tri = trimap.unsqueeze(1).float() * 2 - 1  # sdmatte_nodes.py:351
    flags = is_transparent if isinstance(is_transparent, (list, tuple)) else [is_transparent] * B
    is_trans = torch.tensor([1 if f else 0 for f in flags], device=device)

    aux_latent = vae_encode(c, tri.repeat(1, 3, 1, 1), "enc_tri")  # meta_arch.py:139-145
    assert prompt in ("trimap", "mask", "bbox_mask", "point_mask")
    if coords is None:
        assert prompt == "trimap", "mask / bbox / point prompts need their coordinates"
        coor = torch.tensor([[0.0, 0.0, 1.0, 1.0]] * B, device=device)  # sdmatte_nodes.py:353
    else:
        coor = coords.to(device).float()
    if prompt == "point_mask":
        coor_emb = point_coords_embedding(coor)  # meta_arch.py:153-176
    else:
        assert coor.shape == (B, 4)
        coor_emb = timestep_embedding(coor.flatten(), 320)  # meta_arch.py:181-187
    attention_mask = (tri + 1) / 2  # meta_arch.py:200-204
    attention_mask = F.interpolate(attention_mask, scale_factor=1 / 8, mode="nearest").flatten(start_dim=1)
    rgb_latent = vae_encode(c, rgb, "enc_rgb")  # meta_arch.py:209-212
    ehs = _conv(c, aux_latent, "unet.aux_conv_in")  # meta_arch.py:215-218
    ehs = ehs.view(B, 1024, -1).permute(0, 2, 1)
    trans = 1 - is_trans  # meta_arch.py:237-238
    unet_input = torch.cat([rgb_latent, aux_latent], dim=1)  # meta_arch.py:244
    label_latent = unet_forward(c, unet_input, trans, ehs, coor_emb, attention_mask, point_prompt=(prompt == "point_mask"))
    label_latent = c.r(label_latent / SCALING_FACTOR)  # meta_arch.py:254
    stacked = vae_decode(c, label_latent)
    label_mean = c.r(stacked.mean(dim=1, keepdim=True))  # meta_arch.py:258
    output = torch.clip(label_mean, -1.0, 1.0)
    output = c.r(c.r(output + 1.0) / 2.0)  # meta_arch.py:259-260
    res = {"alpha": output, "label_mean": label_mean, "unet_in": unet_input, "ctx": ehs, "unet_out_scaled": label_latent}
    if capture is not None:
        capture.update(res)
    return res


# ---------------------------------------------------------------------------------------------------------------
# node-level pre/post-processing (sdmatte_nodes.py:339-353, 362-397) — restated; pinned against the lifted
# reference code in tests/test_oracle_cpu.py
# ---------------------------------------------------------------------------------------------------------------
def preprocess(image_bhwc: torch.Tensor, trimap_bhw: torch.Tensor, size: int):
    from torchvision import transforms

    img = transforms.Resize((size, size), antialias=True)(image_bhwc.permute(0, 3, 1, 2).contiguous())
    tri = transforms.Resize((size, size))(trimap_bhw.unsqueeze(1).contiguous())
    return img.permute(0, 2, 3, 1).contiguous(), tri.squeeze(1).contiguous()


def postprocess(pred_alpha_b1rr: torch.Tensor, image: torch.Tensor, trimap: torch.Tensor, output_mode: str, mask_refine: bool,
                trimap_constraint: float):
    from torchvision import transforms

    orig_h, orig_w = image.shape[1], image.shape[2]
    out = transforms.Resize((orig_h, orig_w))(pred_alpha_b1rr)
    out = out.squeeze(1).clamp(0, 1).detach().cpu()
    if mask_refine:
        trimap_cpu = trimap.cpu()
        fg = trimap_cpu > trimap_constraint
        bg = trimap_cpu < (1.0 - trimap_constraint)
        unknown = ~(fg | bg)
        refined = out.clone()
        refined[bg] = 0.0
        refined[fg] = torch.clamp(refined[fg] * 1.2, 0, 1)
        refined[(refined < 0.3) & unknown] = 0.0
        out = refined
    ae = out.unsqueeze(-1)
    if output_mode == "alpha_only":
        matted = torch.zeros_like(image.cpu())
    elif output_mode == "matted_rgba":
        matted = torch.cat([image.cpu(), ae.expand(-1, -1, -1, 1)], dim=-1)
    elif output_mode == "matted_rgb":
        fgm = (trimap.cpu().unsqueeze(-1) > 0.2) & (ae > 0.1)
        matted = image.cpu() * fgm.float()
    else:
        matted = image.cpu() * ae
    return out, matted
