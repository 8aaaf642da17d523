"""Generate the committed golden vectors (run in the BUILD container, where /root/reference exists).

1. alpha_R64_seed1234.npz : oracle (fp32) output for the seeded synthetic checkpoint/input at R=64.  The checkpoint is
   ~3.8 GB and cannot be committed; it is regenerated from its seed (oracle/synth.py) on every machine.
2. ref_lifted.npz         : outputs of the REFERENCE'S OWN functions (lifted from /root/reference with ast, see
   oracle/ref_lifted.py) on seeded inputs: attention-mask preparation and scores (replace.py:20-122), UNet surgery
   (utils.py:13-41), node resize helpers and post-processing (sdmatte_nodes.py:204-214,362-397).  These pin the oracle's
   restatement of those pieces to the reference itself.

Usage: python tests/golden/make_golden.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_lifted, sdmatte_oracle as orc, synth  # noqa: E402


def golden_alpha():
    sd = synth.make_checkpoint(seed=1234)
    image, trimap = synth.make_inputs(1, 64, seed=0)
    out = orc.forward(sd, image, trimap, is_transparent=False)
    np.savez_compressed(os.path.join(HERE, "alpha_R64_seed1234.npz"), alpha=out["alpha"][0, 0].numpy().astype(np.float32),
                        label_mean=out["label_mean"][0, 0].numpy().astype(np.float32), trimap=trimap.numpy(),
                        image_sum=np.float64(image.double().sum().item()), input_seed=0, ckpt_seed=1234)
    print("alpha golden:", out["alpha"].mean().item())


def golden_lifted():
    assert ref_lifted.available(), "run in the build container (needs /root/reference)"
    g = torch.Generator().manual_seed(123)
    res = {}
    # ---- attention mask + scores (replace.py:20-122)
    fns = ref_lifted.attention_fns()
    B, heads, L0, L1, d = 2, 3, 64, 16, 8
    tri_mask = torch.randint(0, 3, (B, L0), generator=g).float() / 2  # values {0, .5, 1} like a trimap at latent size
    add_mask = ((1 - tri_mask) * -10000.0).unsqueeze(1)  # as CustomUNet.forward builds it (replace.py:401-403)
    attn = SimpleNamespace(heads=heads, upcast_attention=False, upcast_softmax=False, scale=d ** -0.5)
    m0 = fns.custom_prepare_attention_mask(attn, add_mask, L0, B)
    m1 = fns.custom_prepare_attention_mask(attn, add_mask, L1, B)
    q = torch.randn(B * heads, L1, d, generator=g)
    k = torch.randn(B * heads, L1, d, generator=g)
    probs = fns.custom_get_attention_scores(attn, q, k, m1)
    probs_nomask = fns.custom_get_attention_scores(attn, q, k, None)
    res.update(att_tri=tri_mask.numpy(), att_m0=m0.numpy(), att_m1=m1.numpy(), att_q=q.numpy(), att_k=k.numpy(),
               att_probs=probs.numpy(), att_probs_nomask=probs_nomask.numpy())
    # ---- UNet surgery (utils.py:13-41)
    sfn = ref_lifted.surgery_fns()
    conv = torch.nn.Conv2d(4, 320, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g))
        conv.bias.copy_(torch.randn(conv.bias.shape, generator=g))
    unet = SimpleNamespace(conv_in=conv, config={})
    w_before = conv.weight.detach().clone()
    sfn.add_aux_conv_in(unet)
    sfn.replace_unet_conv_in(unet, 2)
    res.update(sur_w_before=w_before.numpy(), sur_conv_in_w=unet.conv_in.weight.detach().numpy(),
               sur_aux_w_sum=np.float64(unet.aux_conv_in.weight.detach().double().abs().sum().item()),
               sur_aux_w_head=unet.aux_conv_in.weight.detach()[:320].numpy())
    # ---- node helpers (sdmatte_nodes.py:204-214) and post-processing (sdmatte_nodes.py:362-397)
    nfn = ref_lifted.node_helpers()
    img = torch.rand(1, 3, 50, 70, generator=g)
    msk = (torch.rand(1, 1, 50, 70, generator=g) > 0.5).float()
    res.update(node_img=img.numpy(), node_msk=msk.numpy(), node_img_r=nfn._resize_norm_image_bchw(img, (32, 32)).numpy(),
               node_msk_r=nfn._resize_mask_b1hw(msk, (32, 32)).numpy())
    image = torch.rand(2, 40, 30, 3, generator=g)
    trimap = (torch.randint(0, 3, (2, 40, 30), generator=g).float() / 2)
    pred = torch.rand(2, 1, 32, 32, generator=g)
    for mode in ("alpha_only", "matted_rgba", "matted_rgb"):
        for refine in (True, False):
            o, m = ref_lifted.node_postprocess(pred, image, trimap, mode, refine, 0.8)
            res[f"post_{mode}_{int(refine)}_alpha"] = o.numpy()
            res[f"post_{mode}_{int(refine)}_matted"] = m.numpy()
    res.update(post_image=image.numpy(), post_trimap=trimap.numpy(), post_pred=pred.numpy())
    np.savez_compressed(os.path.join(HERE, "ref_lifted.npz"), **res)
    print("lifted goldens:", len(res), "arrays")


def golden_prepost():
    """3. prepost_lifted.npz : the reference's resize helpers (sdmatte_nodes.py:204-214) and post-processing statements
    (sdmatte_nodes.py:362-397) run on NON-square, non-R inputs with an fp16 model output (as on the reference's CUDA path) —
    the known answers for the GPU pre/post kernels (csrc/prepost.cu)."""
    assert ref_lifted.available(), "run in the build container (needs /root/reference)"
    g = torch.Generator().manual_seed(321)
    nfn = ref_lifted.node_helpers()
    res = {}
    R = 32
    for tag, (H, W) in {"down": (60, 41), "up": (20, 28), "mixed": (24, 90)}.items():
        image = torch.rand(2, H, W, 3, generator=g)
        # trimap with soft edges (values other than 0 / .5 / 1 exercise the thresholds)
        trimap = (torch.randint(0, 3, (2, H, W), generator=g).float() / 2)
        trimap = torch.where(torch.rand(2, H, W, generator=g) < 0.1, torch.rand(2, H, W, generator=g), trimap)
        img_r = nfn._resize_norm_image_bchw(image.permute(0, 3, 1, 2).contiguous(), (R, R))   # normalised (x-0.5)/0.5
        tri_r = nfn._resize_mask_b1hw(trimap.unsqueeze(1).contiguous(), (R, R))
        pred = (torch.rand(2, 1, R, R, generator=g) * 1.2 - 0.1).half()    # fp16, slightly outside [0,1] -> clamp matters
        res.update({f"{tag}_image": image.numpy(), f"{tag}_trimap": trimap.numpy(), f"{tag}_img_r": img_r.numpy(),
                    f"{tag}_tri_r": tri_r.numpy(), f"{tag}_pred": pred.numpy()})
        for mode in ("alpha_only", "matted_rgba", "matted_rgb"):
            for refine, c in ((True, 0.8), (True, 0.3), (False, 0.8)):
                o, m = ref_lifted.node_postprocess(pred, image, trimap, mode, refine, c)
                assert o.dtype == torch.float16
                key = f"{tag}_{mode}_{int(refine)}_{c}"
                res[key + "_alpha"] = o.numpy()
                if mode != "alpha_only":
                    if refine and c == 0.8:  # keep the fixture small: composition is checked for one refine setting per mode
                        res[key + "_matted"] = m.numpy()
                else:
                    assert float(m.abs().max()) == 0.0 and m.shape == image.shape
    np.savez_compressed(os.path.join(HERE, "prepost_lifted.npz"), **res)
    print("prepost goldens:", len(res), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["lifted", "alpha", "prepost"]
    if "lifted" in which:
        golden_lifted()
    if "alpha" in which:
        golden_alpha()
    if "prepost" in which:
        golden_prepost()
