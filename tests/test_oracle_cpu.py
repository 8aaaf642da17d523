"""CPU suite (no GPU): the oracle against the committed golden vectors (incl. the ones produced by the reference's own
functions), host-side logic, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lifted():
    return np.load(os.path.join(GOLD, "ref_lifted.npz"))


def test_parameter_counts():
    """Architecture self-check without diffusers: published SD-2.1 UNet = 865.91 M, SD VAE = 83.65 M parameters."""
    from oracle import synth

    s = synth.param_shapes(include_unused=True)
    unet = sum(int(np.prod(v)) for k, v in s.items() if k.startswith("unet."))
    vae = sum(int(np.prod(v)) for k, v in s.items() if k.startswith("vae."))
    assert vae == 83_653_863
    assert unet == 873_030_852  # CustomUNet: + 8-ch conv_in, aux_conv_in, point/bbox embeddings
    extra = (320 * 4 * 9) + (1024 * 4 * 9 + 1024) + (1280 * 1680 + 1280 + 1280 * 1280 + 1280) + (1280 * 1280 + 1280) * 2
    assert unet - extra == 865_910_724
    assert s["unet.point_embedding.linear_1.weight"] == (1280, 1680)  # pinned by the comment at meta_arch.py:108


def test_key_bias_matches_reference_mask_functions(lifted):
    """oracle.prepare_key_bias == custom_prepare_attention_mask (replace.py:20-72), executed from /root/reference."""
    from oracle import sdmatte_oracle as orc

    tri = torch.from_numpy(lifted["att_tri"])
    add = ((1 - tri) * -10000.0).unsqueeze(1)
    heads = 3
    m0 = orc.prepare_key_bias(add, 64).repeat_interleave(heads, dim=0)
    m1 = orc.prepare_key_bias(add, 16).repeat_interleave(heads, dim=0)
    assert np.array_equal(m0.numpy(), lifted["att_m0"])
    assert np.array_equal(m1.numpy(), lifted["att_m1"])
    # the two nearest resizes compose to strided sampling (what the CUDA key_bias kernel does)
    assert torch.equal(orc.prepare_key_bias(add, 16), add.view(2, 1, 8, 8)[:, :, ::2, ::2].reshape(2, 1, 16))


def test_attention_scores_match_reference(lifted):
    """softmax(scale q k^T + bias) as the oracle computes it == custom_get_attention_scores (replace.py:75-122)."""
    q, k = torch.from_numpy(lifted["att_q"]), torch.from_numpy(lifted["att_k"])
    m1 = torch.from_numpy(lifted["att_m1"])
    scale = 8 ** -0.5
    probs = (torch.matmul(q, k.transpose(-1, -2)) * scale + m1).softmax(-1)
    np.testing.assert_allclose(probs.numpy(), lifted["att_probs"], rtol=1e-5, atol=1e-7)
    probs0 = (torch.matmul(q, k.transpose(-1, -2)) * scale).softmax(-1)
    np.testing.assert_allclose(probs0.numpy(), lifted["att_probs_nomask"], rtol=1e-5, atol=1e-7)


def test_key_compaction_drops_only_exact_zero_probabilities(lifted):
    """The engine's attn1 streams only the keys with bias >= max(bias) - 2500 (key_compact_kernel).  Against the output of the
    reference's OWN custom_get_attention_scores (golden `att_probs`, replace.py:75-122): every dropped key has probability
    exactly 0.0 there, and the softmax over the kept keys alone reproduces the reference's probabilities of those keys."""
    q, k = torch.from_numpy(lifted["att_q"]), torch.from_numpy(lifted["att_k"])
    m1 = torch.from_numpy(lifted["att_m1"])                  # (B*heads, 1, L) additive mask: 0 / -5000 / -10000
    probs = torch.from_numpy(lifted["att_probs"])            # (B*heads, Lq, L)
    keep = m1 >= m1.max(dim=-1, keepdim=True).values - 2500.0
    assert 0 < keep.float().mean().item() < 1                # the golden trimap has foreground AND non-foreground keys
    assert (probs[(~keep).expand_as(probs)] == 0).all(), "a dropped key has a non-zero probability in the reference"
    scale = 8 ** -0.5
    for i in range(q.shape[0]):
        kk = keep[i, 0]
        sub = (torch.matmul(q[i], k[i][kk].t()) * scale + m1[i, :, kk]).softmax(-1)
        np.testing.assert_allclose(sub.numpy(), probs[i][:, kk].numpy(), rtol=1e-5, atol=1e-7)


def test_unet_surgery_semantics(lifted):
    """replace_unet_conv_in / add_aux_conv_in (utils.py:13-41): conv_in becomes 8-ch = tiled weights / 2; aux_conv_in 4->1024."""
    w = lifted["sur_w_before"]
    np.testing.assert_allclose(lifted["sur_conv_in_w"], np.concatenate([w, w], axis=1) / 2, rtol=0, atol=0)
    np.testing.assert_array_equal(lifted["sur_aux_w_head"], w)
    from oracle import synth

    s = synth.param_shapes()
    assert s["unet.conv_in.weight"] == (320, 8, 3, 3) and s["unet.aux_conv_in.weight"] == (1024, 4, 3, 3)


def test_preprocess_matches_reference_helpers(lifted):
    """oracle.preprocess + Normalize == _resize_norm_image_bchw / _resize_mask_b1hw (sdmatte_nodes.py:204-214)."""
    from oracle import sdmatte_oracle as orc

    img = torch.from_numpy(lifted["node_img"]).permute(0, 2, 3, 1)
    msk = torch.from_numpy(lifted["node_msk"]).squeeze(1)
    i2, m2 = orc.preprocess(img, msk, 32)
    np.testing.assert_allclose(((i2.permute(0, 3, 1, 2) - 0.5) / 0.5).numpy(), lifted["node_img_r"], atol=1e-6)
    np.testing.assert_allclose(m2.unsqueeze(1).numpy(), lifted["node_msk_r"], atol=1e-6)


@pytest.mark.parametrize("mode", ["alpha_only", "matted_rgba", "matted_rgb"])
@pytest.mark.parametrize("refine", [True, False])
def test_postprocess_matches_reference(lifted, mode, refine):
    """oracle.postprocess == the reference's post-processing statements (sdmatte_nodes.py:362-397) run verbatim."""
    from oracle import sdmatte_oracle as orc

    image, trimap, pred = (torch.from_numpy(lifted[k]) for k in ("post_image", "post_trimap", "post_pred"))
    a, m = orc.postprocess(pred, image, trimap, mode, refine, 0.8)
    np.testing.assert_array_equal(a.numpy(), lifted[f"post_{mode}_{int(refine)}_alpha"])
    np.testing.assert_array_equal(m.numpy(), lifted[f"post_{mode}_{int(refine)}_matted"])


def test_node_runs_pre_and_post_on_the_device():
    """§8(f) n1: the node has no torch/CPU resize or refine code of its own — everything between its argument checks and its
    return is the C-ABI node call (sdm_apply_matte_host: staging, resize, forward, post-processing; GPU parity of the pre/post
    kernels in tests/test_prepost_gpu.py, of the whole call in tests/test_multigpu_gpu.py)."""
    import inspect

    import __graft_entry__ as ge

    nodes = ge.load_package().sdmatte_nodes
    src = inspect.getsource(nodes.SDMatteApply.apply_matte)
    assert "eng.apply_host(" in src and ".cuda(" not in src and ".to(device=device" not in src
    assert "interpolate(" not in inspect.getsource(nodes) and ".clamp(" not in src and "refined" not in src
    node = nodes.SDMatteApply()
    assert node.RETURN_TYPES == ("MASK", "IMAGE") and node.FUNCTION == "apply_matte"


def test_timestep_embedding_formula():
    from oracle import sdmatte_oracle as orc

    e = orc.timestep_embedding(torch.tensor([0.0, 1.0]), 320)
    assert e.shape == (2, 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))
    assert abs(e[1, 0].item() - math.cos(1.0)) < 1e-6 and abs(e[1, 160].item() - math.sin(1.0)) < 1e-6
    assert abs(e[1, 159].item() - math.cos(math.exp(-math.log(10000) * 159 / 160))) < 1e-6


def test_oracle_reproduces_committed_golden_alpha():
    """The oracle regenerates the committed golden (checkpoint and inputs from their seeds) — guards against drift of the
    generators and of torch CPU kernels between the build container and the GPU box."""
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    g = np.load(os.path.join(GOLD, "alpha_R64_seed1234.npz"))
    image, trimap = synth.make_inputs(1, 64, seed=int(g["input_seed"]))
    assert np.array_equal(trimap.numpy(), g["trimap"])
    assert abs(image.double().sum().item() - float(g["image_sum"])) < 1e-3
    sd = synth.make_checkpoint(seed=int(g["ckpt_seed"]))
    out = orc.forward(sd, image, trimap)
    d = np.abs(out["alpha"][0, 0].numpy() - g["alpha"]).max()
    assert d < 2e-4, d
    out16 = orc.forward(sd, image, trimap, mode="fp16sim")
    d16 = (out16["alpha"] - out["alpha"]).abs().max().item()
    assert d16 < 1e-2  # fp16 rounding points move alpha by a few fp16 ulps at most
    # sliced attention (the reference's CUDA path uses SlicedAttnProcessor(slice_size=1)) is numerically the same op
    outs = orc.forward(sd, image, trimap, sliced=True)
    assert (outs["alpha"] - out["alpha"]).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------------ boundary
def test_cabi_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    pkg = ge.load_package()
    header = open(os.path.join(ROOT, "include", "sdmatte_b200.h")).read()
    declared = set(re.findall(r"\b(sdm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(pkg.engine.lib_path())
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/sdmatte_b200.h but not exported"
    assert declared == set(pkg.engine.EXPORTS)
    lib.sdm_version.restype = ctypes.c_int
    assert lib.sdm_version() >= 100


def test_conv_variant_is_a_function_of_geometry(pkg):
    """The engine's choice between the tap-per-box, resident-halo and swapped-operand 3x3 kernels (they differ in K order and in
    the GroupNorm-partials grouping, i.e. in output bits) is host logic with NO batch size in its signature, so a sample gives
    the same bits alone and in a batch.  Pin the choice for every 3x3 conv geometry of the model at all node resolutions."""
    E = pkg.engine
    for R in (512, 640, 768, 896, 1024):
        S = R // 8
        # VAE: 128 ch @R, 256 @R/2, 512 @R/4 and R/8 (encoder and decoder); decoder upsample convs at 2x of the 512 / 512 / 256 levels
        assert E.conv_variant(3, 1, 128, R, R, has_res=1) == 3, R          # swapped operands (channels on M)
        assert E.conv_variant(3, 1, 128, R, R) == 3
        assert E.conv_variant(3, 1, 256, R // 2, R // 2, has_res=1) == 1   # resident halo, 256-wide tiles
        assert E.conv_variant(3, 1, 512, R // 4, R // 4) == 1
        assert E.conv_variant(3, 1, 512, R // 8, R // 8, ups2=1) == 1
        assert E.conv_variant(3, 1, 256, R, R) == 1
        # UNet: 320 @S, 640 @S/2 (160-wide tiles), 1280 @S/4
        assert E.conv_variant(3, 1, 320, S, S) == 2
        assert E.conv_variant(3, 1, 640, S // 2, S // 2, has_res=1) == 2
        # 1280 ch @S/4, S/8: halo where the 8 x 16 patches tile the grid like the default patch does (same GroupNorm slot count)
        assert E.conv_variant(3, 1, 1280, S // 4, S // 4) in (0, 1)
        assert E.conv_variant(3, 1, 1280, S // 8, S // 8) == (1 if S // 8 == 16 else 0)  # below 16 rows: no halo patch
        # everything else keeps one TMA box per tap
        assert E.conv_variant(3, 2, 320, S, S) == 0       # Downsample2D
        assert E.conv_variant(1, 1, 256, R // 2, R // 2) == 0   # 1x1 shortcut
        assert E.conv_variant(3, 1, 8, S, S) == 0         # skinny outputs
        assert E.conv_variant(3, 1, 256, R // 2, R // 2, mode=3) == 0  # only the fp16 epilogue has the variants
    assert E.conv_variant(3, 1, 1280, 32, 32) == 1 and E.conv_variant(3, 1, 1280, 20, 20) == 0  # R = 1024 / 640 at S/4
    for hw in (8, 12):  # tiny latents (tests at R = 64 / 96): no variant kernel
        assert E.conv_variant(3, 1, 1280, hw, hw) == 0 and E.conv_variant(3, 1, 128, hw, hw) == 0


def test_conv_variant_keeps_the_groupnorm_slot_count(pkg):
    """The conv epilogues write one GroupNorm-partials slot per 128 output pixels into a buffer sized by
    sdm_k_conv_tiles_per_image(H, W) (default patch).  The halo kernel (8 x 16 patches) and the swapped-operand kernel (16 x 16
    patches, two slots each) must cover exactly that many slots wherever the engine selects them."""
    E = pkg.engine
    checked = 0
    for H in range(8, 1025, 8):
        for W in {H, max(8, H // 2), min(1024, H + 24)}:
            slots = E.conv_tiles_per_image(H, W)
            for N in (128, 256, 320, 512, 640, 1280):
                v = E.conv_variant(3, 1, N, H, W)
                if v == 3:
                    assert 2 * (-(-W // 16)) * (-(-H // 16)) == slots, (H, W, N)
                    checked += 1
                elif v in (1, 2):
                    assert (-(-W // 8)) * (-(-H // 16)) == slots, (H, W, N)
                    checked += 1
    assert checked > 500


def test_no_cpu_fallback_and_no_oracle_on_product_path():
    import __graft_entry__ as ge

    pkg = ge.load_package()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pkg.engine.Engine(0)
    pkg_dir = os.path.join(ROOT, "comfyui-sdmatte_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"


def test_node_surface_matches_reference_schema():
    """Widget schema / return types of the reference node (sdmatte_nodes.py:219-255,408-414)."""
    import __graft_entry__ as ge

    pkg = ge.load_package()
    assert set(pkg.NODE_CLASS_MAPPINGS) == {"SDMatteApply"}
    assert pkg.NODE_DISPLAY_NAME_MAPPINGS == {"SDMatteApply": "Apply SDMatte"}
    cls = pkg.NODE_CLASS_MAPPINGS["SDMatteApply"]
    it = cls.INPUT_TYPES()
    req = it["required"]
    assert list(req) == ["ckpt_name", "image", "trimap", "inference_size", "is_transparent", "output_mode", "mask_refine", "trimap_constraint"]
    assert req["ckpt_name"][0] == ["SDMatte.safetensors", "SDMatte_plus.safetensors"]
    assert req["image"][0] == "IMAGE" and req["trimap"][0] == "MASK"
    assert req["inference_size"][0] == [512, 640, 768, 896, 1024] and req["inference_size"][1]["default"] == 1024
    assert req["is_transparent"][1]["default"] is False and req["mask_refine"][1]["default"] is True
    assert req["output_mode"][0] == ["alpha_only", "matted_rgba", "matted_rgb"]
    tc = req["trimap_constraint"][1]
    assert (tc["default"], tc["min"], tc["max"], tc["step"]) == (0.8, 0.1, 1.0, 0.1)
    assert it["optional"]["force_cpu"][1]["default"] is False
    assert cls.RETURN_TYPES == ("MASK", "IMAGE") and cls.RETURN_NAMES == ("alpha_mask", "matted_image")
    assert cls.FUNCTION == "apply_matte" and cls.CATEGORY == "Matting/SDMatte"
    import inspect

    assert list(inspect.signature(cls.apply_matte).parameters) == ["self", "ckpt_name", "image", "trimap", "inference_size", "is_transparent",
                                                                    "output_mode", "mask_refine", "trimap_constraint", "force_cpu"]
    nodes = pkg.sdmatte_nodes
    os.environ["SDMATTE_OFFLINE"] = "1"  # no network here: a missing checkpoint must be a clear error, not a download attempt
    try:
        with pytest.raises(ValueError):
            nodes.find_checkpoint("nope.safetensors")
        with pytest.raises(FileNotFoundError):
            nodes.find_checkpoint("SDMatte.safetensors")
    finally:
        del os.environ["SDMATTE_OFFLINE"]


def test_flop_model_matches_survey():
    from bench import TFLOP_PER_MATTE

    assert TFLOP_PER_MATTE[1024] == 28.785 and TFLOP_PER_MATTE[512] == 5.952


def test_point_coordinate_embedding_padding_rule():
    """meta_arch.py:153-176: N coordinates are zero-padded to the first i >= N dividing 1680, 1680 / i sinusoid channels each
    (an odd channel count gets one zero column, diffusers get_timestep_embedding), always 1680 values per sample."""
    from oracle import sdmatte_oracle as orc

    for n, i in ((1, 1), (4, 4), (11, 12), (13, 14), (16, 16), (20, 20), (41, 42), (100, 105)):
        c = torch.rand(3, n)
        e = orc.point_coords_embedding(c)
        assert e.shape == (3, 1680)
        dim = 1680 // i
        assert 1680 % i == 0 and all(1680 % j for j in range(n, i))
        per = e.view(3, i, dim)
        half = dim // 2
        # coordinate 0, frequency 0: cos(t), sin(t); padded coordinates embed t = 0: cos = 1, sin = 0
        assert torch.allclose(per[:, 0, 0], torch.cos(c[:, 0])) and torch.allclose(per[:, 0, half], torch.sin(c[:, 0]))
        if i > n:
            assert torch.allclose(per[:, n:, :half], torch.ones(3, i - n, half)) and torch.allclose(per[:, n:, half:2 * half], torch.zeros(3, i - n, half))
        if dim % 2:
            assert (per[:, :, -1] == 0).all()
