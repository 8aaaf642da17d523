"""End-to-end parity (GPU): the sm_100a engine (through the C ABI) against the CPU oracle on the same seeded inputs
and the same synthetic checkpoint, plus size-independent properties at the benchmark's full size.

Tolerance: north_star asks |d_alpha| <= 1e-3 (fp16 grid near alpha=1 is 4.9e-4, i.e. ~2 ulp).  The engine stores
activations in fp16 (like the reference's autocast path) while the oracle's default mode is fp32, so the bound used
against the fp32 oracle is 4e-3 max / 5e-4 mean; against the oracle's fp16-rounding-point emulation ("fp16sim") the
two differ only by accumulation order and exp/erf implementations.  Measured values are printed and recorded in
gpurun_out/parity.json for DESIGN.md.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ckpt():
    from oracle import synth

    return synth.make_checkpoint(seed=1234)


@pytest.fixture(scope="module")
def engine(pkg, ckpt):
    eng = pkg.engine.Engine(0)
    used, unexpected = eng.load_state_dict(ckpt)
    assert used == len(ckpt) and unexpected == 0
    yield eng
    eng.close()


def _record(name, **kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "parity.json")
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
    data[name] = kw
    json.dump(data, open(path, "w"), indent=1)


@pytest.mark.parametrize("R,B", [(64, 1), (128, 2), (192, 1), (256, 1)])  # 192: latent 24 -> 12, 6, 3 (ragged tile patches)
def test_alpha_matches_oracle(engine, ckpt, R, B):
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    image, trimap = synth.make_inputs(B, R, seed=R)
    flags = [bool(i % 2) for i in range(B)]
    alpha, pre = engine.forward(image.cuda(), trimap.cuda(), flags, want_premean=True)
    torch.cuda.synchronize()
    ref = orc.forward(ckpt, image, trimap, is_transparent=flags)
    a = alpha.float().cpu()
    for name in ("unet_in", "ctx", "unet_out_scaled"):
        got = engine.debug_tensor(name).float().cpu()
        want = ref[name]
        want = want.permute(0, 2, 3, 1) if name != "ctx" else want.reshape(got.shape)
        rel = ((got - want).abs().max() / want.abs().max()).item()
        print(f"[tap] {name}: max rel-to-range err {rel:.3e}")
        assert rel < 6e-3, f"{name} diverges: {rel}"  # measured 1.6e-3 .. 2.8e-3 (profiles/r2i_parity.json: error_growth, max/range column)
    d = (a - ref["alpha"].squeeze(1)).abs()
    dm = (pre.float().cpu() - ref["label_mean"].squeeze(1)).abs()
    print(f"[parity fp32-oracle] R={R} B={B} max|da|={d.max():.3e} mean|da|={d.mean():.3e} max|dmean|={dm.max():.3e}")
    _record(f"alpha_fp32_R{R}_B{B}", max_abs=d.max().item(), mean_abs=d.mean().item(), premean_max_abs=dm.max().item())
    assert d.max().item() <= 4e-3 and d.mean().item() <= 5e-4


def test_alpha_matches_fp16sim_oracle(engine, ckpt):
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    R, B = 128, 1
    image, trimap = synth.make_inputs(B, R, seed=5)
    alpha = engine.forward(image.cuda(), trimap.cuda(), False).float().cpu()
    ref = orc.forward(ckpt, image, trimap, is_transparent=False, mode="fp16sim")
    d = (alpha - ref["alpha"].squeeze(1)).abs()
    print(f"[parity fp16sim-oracle] max|da|={d.max():.3e} mean|da|={d.mean():.3e}")
    _record("alpha_fp16sim_R128", max_abs=d.max().item(), mean_abs=d.mean().item())
    assert d.max().item() <= 4e-3 and d.mean().item() <= 5e-4


def test_golden_alpha(engine):
    """Committed golden vector produced by the oracle in the build container (tests/golden/make_golden.py)."""
    import numpy as np
    from oracle import synth

    path = os.path.join(ROOT, "tests", "golden", "alpha_R64_seed1234.npz")
    g = np.load(path)
    image, trimap = synth.make_inputs(1, 64, seed=int(g["input_seed"]))
    assert np.array_equal(trimap.numpy(), g["trimap"]), "input generator drifted from the committed golden"
    alpha = engine.forward(image.cuda(), trimap.cuda(), False).float().cpu().numpy()
    d = np.abs(alpha[0] - g["alpha"])
    print(f"[golden] max|da|={d.max():.3e}")
    assert d.max() <= 4e-3


def test_batch_independence_and_determinism(engine):
    """Samples are independent (per-sample GroupNorm/attention): element i of a batch == the same input run alone, bit for bit;
    and two runs of the same batch are bit-identical (no atomics on the path)."""
    from oracle import synth

    R, B = 128, 3
    image, trimap = synth.make_inputs(B, R, seed=11)
    img, tri = image.cuda(), trimap.cuda()
    flags = [False, True, False]
    a1 = engine.forward(img, tri, flags).clone()
    a2 = engine.forward(img, tri, flags).clone()
    assert torch.equal(a1, a2)
    for i in range(B):
        ai = engine.forward(img[i:i + 1].contiguous(), tri[i:i + 1].contiguous(), flags[i]).clone()
        assert torch.equal(ai[0], a1[i]), f"sample {i} depends on its batch"


def test_transparent_flag_changes_output(engine):
    from oracle import synth

    image, trimap = synth.make_inputs(1, 64, seed=3)
    a0 = engine.forward(image.cuda(), trimap.cuda(), False).clone()
    a1 = engine.forward(image.cuda(), trimap.cuda(), True).clone()
    assert not torch.equal(a0, a1)


def test_forward_host_equals_forward(engine):
    from oracle import synth

    image, trimap = synth.make_inputs(2, 64, seed=4)
    a_dev = engine.forward(image.cuda(), trimap.cuda(), False).cpu()
    a_host = engine.forward_host(image.pin_memory(), trimap.pin_memory(), False)
    assert torch.equal(a_dev, a_host)


def test_all_background_trimap_edge(engine, ckpt):
    """No foreground key anywhere: every key bias is -10000 and the softmax renormalises."""
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    image, _ = synth.make_inputs(1, 64, seed=9)
    trimap = torch.zeros(1, 64, 64)
    alpha = engine.forward(image.cuda(), trimap.cuda(), False).float().cpu()
    ref = orc.forward(ckpt, image, trimap)
    d = (alpha - ref["alpha"].squeeze(1)).abs()
    print(f"[all-bg] max|da|={d.max():.3e}")
    assert d.max().item() <= 6e-3


def test_node_end_to_end_with_resize(pkg, ckpt, engine):
    """The ComfyUI node surface: non-square input, resize to 128, mask_refine + matted_rgb, compared with the oracle pipeline."""
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    nodes = pkg.sdmatte_nodes
    nodes.register_state_dict("SDMatte.safetensors", ckpt)
    node = nodes.SDMatteApply()
    image, trimap = synth.make_inputs(1, 128, seed=21, Hin=150, Win=100)
    alpha, matted = node.apply_matte("SDMatte.safetensors", image, trimap, 128, False, "matted_rgb", True, 0.8)
    assert alpha.shape == (1, 150, 100) and matted.shape == (1, 150, 100, 3)
    img_r, tri_r = orc.preprocess(image, trimap, 128)
    ref = orc.forward(ckpt, img_r, tri_r)
    ref_alpha, ref_matted = orc.postprocess(ref["alpha"], image, trimap, "matted_rgb", True, 0.8)
    d = (alpha.float() - ref_alpha.float()).abs()
    # thresholded refinement can flip pixels that sit within fp16 noise of 0.3 / 1/1.2: compare away from those
    stable = ((ref_alpha - 0.3).abs() > 0.01) | (ref_alpha == 0)
    print(f"[node] max|da| (stable px)={d[stable].max():.3e}, unstable px={int((~stable).sum())}")
    assert d[stable].max().item() <= 6e-3
    with pytest.raises(RuntimeError):
        node.apply_matte("SDMatte.safetensors", image, trimap, 128, False, "alpha_only", True, 0.8, force_cpu=True)
    with pytest.raises(ValueError):
        node.apply_matte("SDMatte.safetensors", torch.rand(1, 64, 64, 4), trimap, 128, False, "alpha_only", True, 0.8)


@pytest.mark.parametrize("R", [640, 896, 1024])
def test_full_size_properties(engine, R):
    """BASELINE sizes (inference_size 640 / 896 have non-power-of-two latent grids 80/112; 1024 is the headline): alpha in [0,1] on the fp16 grid, finite, deterministic, and batch element == single run."""
    from oracle import synth

    B = 2
    image, trimap = synth.make_inputs(B, R, seed=2)
    img, tri = image.cuda(), trimap.cuda()
    a = engine.forward(img, tri, False).clone()
    assert torch.isfinite(a.float()).all() and a.min() >= 0 and a.max() <= 1
    frac_sat = ((a == 0) | (a == 1)).float().mean().item()
    print(f"[1024] saturated fraction {frac_sat:.3f}, mean alpha {a.float().mean():.3f}")
    assert frac_sat < 0.5
    a_single = engine.forward(img[1:2].contiguous(), tri[1:2].contiguous(), False)
    assert torch.equal(a_single[0], a[1])


@pytest.mark.parametrize("R,B", [(128, 2), (512, 1)])
def test_key_compaction_is_transparent(pkg, ckpt, engine, R, B):
    """attn1 streams only the keys whose softmax probability can be non-zero (key_compact_kernel).  The dropped keys have
    probability exactly 0 in the reference (kernel-level proof: test_attention_compacted_keys_matches_full), so switching the
    compaction off only changes the summation order of the kept keys (tile boundaries move).  One flipped fp16 rounding early
    in the UNet re-draws the whole rounding-noise realisation downstream, so the two alphas differ like two fp16 runs do
    (measured r1q: max 2.4e-3 / 2.9e-3 at R = 128 / 512): the bound is the engine-vs-oracle one, and both runs must meet it."""
    from oracle import synth

    image, trimap = synth.make_inputs(B, R, seed=31)
    img, tri = image.cuda(), trimap.cuda()
    a_on = engine.forward(img, tri, False).clone()
    os.environ["SDM_ATTN_COMPACT"] = "0"
    try:
        eng2 = pkg.engine.Engine(0)
        eng2.load_state_dict(ckpt)
        a_off = eng2.forward(img, tri, False).clone()
        n_off = eng2.stats()["launches"]
        eng2.close()
    finally:
        del os.environ["SDM_ATTN_COMPACT"]
    n_on = engine.stats()["launches"]
    d = (a_on.float() - a_off.float()).abs()
    print(f"[compact] R={R} launches on/off {n_on}/{n_off} max|da|={d.max():.3e} mean|da|={d.mean():.3e} differing px {(d > 0).float().mean():.4f}")
    _record(f"compact_on_off_R{R}_B{B}", max_abs=d.max().item(), mean_abs=d.mean().item())
    assert n_on == n_off + 17  # 16 gathers + the compaction kernel
    assert d.max().item() <= 4e-3 and d.mean().item() <= 5e-4


def test_native_safetensors_load_equals_state_dict_load(pkg, ckpt, tmp_path):
    """the engine loaded through the native .safetensors reader (header parse + mmap, repacked straight from the
    mapping) produces the same bits as the engine loaded from the same tensors passed as a state dict."""
    safetensors_torch = pytest.importorskip("safetensors.torch")
    from oracle import synth

    sd16 = {k: v.half().contiguous() for k, v in ckpt.items()}  # fp16 file: half the bytes on disk, same path through the loader
    path = str(tmp_path / "SDMatte_synth.safetensors")
    safetensors_torch.save_file(sd16, path)
    image, trimap = synth.make_inputs(1, 64, seed=17)
    a = pkg.engine.Engine(0)
    used_a, _ = a.load_state_dict(sd16)
    alpha_a = a.forward(image.cuda(), trimap.cuda(), False).clone()
    a.close()
    b = pkg.engine.Engine(0)
    used_b, unexpected_b = b.load_safetensors(path)
    alpha_b = b.forward(image.cuda(), trimap.cuda(), False).clone()
    b.close()
    os.remove(path)
    assert used_a == used_b and unexpected_b == 0
    assert torch.equal(alpha_a, alpha_b)
