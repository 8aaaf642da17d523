"""SURVEY 8(f) n3 — the SDMatte model's other visual prompts (mask / bbox mask with 4 box coordinates through bbox_embedding, point
mask with N point coordinates through point_embedding; SDMatte.forward meta_arch.py:130-197, CustomUNet.forward replace.py:446-455)
through sdm_forward_prompt, against the oracle run with the same prompt on the GPU (fp32, TF32 off).  The node's own trimap prompt
keeps its load-time-folded constants; the per-sample chain (csrc/cond_embed.cu) must reproduce them for coords = [0,0,1,1]."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ckpt():
    from oracle import synth

    return synth.make_checkpoint(seed=1234, include_unused=True)  # with unet.point_embedding.*


@pytest.fixture(scope="module")
def sd_gpu(ckpt):
    from oracle import sdmatte_oracle as orc

    return orc.to_device(ckpt, "cuda")


@pytest.fixture(scope="module")
def engine(pkg, ckpt):
    eng = pkg.engine.Engine(0)
    used, unexpected = eng.load_state_dict(ckpt)
    assert unexpected == 0 and used == len(ckpt)  # the point_embedding tensors are consumed now
    yield eng
    eng.close()


def _d(a, b):
    d = (a.float() - b.float()).abs()
    return d.max().item(), d.mean().item()


def test_default_box_reproduces_the_folded_trimap_path(engine):
    from oracle import synth

    R, B = 128, 2
    image, trimap = synth.make_inputs(B, R, seed=3)
    img, tri = image.cuda(), trimap.cuda()
    flags = [False, True]
    a_fold = engine.forward(img, tri, flags).clone()
    a_dyn = engine.forward_prompt(img, tri, "trimap", torch.tensor([[0.0, 0.0, 1.0, 1.0]] * B), flags)
    mx, mean = _d(a_fold, a_dyn)
    print(f"[prompt] folded vs per-sample chain at [0,0,1,1]: max {mx:.3e} mean {mean:.3e}")
    assert mx <= 4e-3 and mean <= 5e-4  # fp64 host folding vs fp32 GEMVs + fp16 time_emb_proj: fp16-noise apart, not bit-identical


@pytest.mark.parametrize("prompt,ncoords", [("bbox_mask", 4), ("mask", 4), ("point_mask", 20), ("point_mask", 11), ("point_mask", 1)])
def test_prompts_match_oracle(engine, sd_gpu, prompt, ncoords):
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    R, B = 128, 2
    image, aux = synth.make_inputs(B, R, seed=40 + ncoords)
    g = torch.Generator().manual_seed(ncoords)
    coords = torch.rand(B, ncoords, generator=g)
    if ncoords == 4:
        coords = torch.stack([coords[:, 0] * 0.4, coords[:, 1] * 0.4, 0.6 + coords[:, 2] * 0.4, 0.6 + coords[:, 3] * 0.4], dim=1)
    flags = [True, False]
    alpha, pre = engine.forward_prompt(image.cuda(), aux.cuda(), prompt, coords, flags, want_premean=True)
    ref = orc.forward(sd_gpu, image, aux, is_transparent=flags, prompt=prompt, coords=coords, device="cuda")
    mx, mean = _d(alpha, ref["alpha"].squeeze(1))
    base = orc.forward(sd_gpu, image, aux, is_transparent=flags, device="cuda")["alpha"].squeeze(1)
    moved = (ref["alpha"].squeeze(1) - base).abs().max().item()
    print(f"[prompt {prompt} n={ncoords}] engine vs oracle max {mx:.3e} mean {mean:.3e}; the prompt moves alpha by up to {moved:.3e} vs the default box")
    assert mx <= 4e-3 and mean <= 5e-4
    assert moved > 4 * mx, "the test would not notice a wrong embedding"
    # per-sample conditioning: a sample alone gives the same bits as inside the batch
    a1 = engine.forward_prompt(image[1:].cuda(), aux[1:].cuda(), prompt, coords[1:], flags[1:])
    assert torch.equal(a1[0], alpha[1])


def test_point_prompt_needs_point_embedding(pkg):
    from oracle import synth

    eng = pkg.engine.Engine(0)
    eng.load_state_dict(synth.make_checkpoint(seed=1234))  # no unet.point_embedding.*
    image, aux = synth.make_inputs(1, 64, seed=1)
    try:
        with pytest.raises(RuntimeError, match="point_embedding"):
            eng.forward_prompt(image.cuda(), aux.cuda(), "point_mask", torch.rand(1, 5))
        with pytest.raises(RuntimeError, match="4 coordinates"):
            eng.forward_prompt(image.cuda(), aux.cuda(), "bbox_mask", torch.rand(1, 5))
        eng.forward_prompt(image.cuda(), aux.cuda(), "bbox_mask", torch.rand(1, 4))  # bbox prompts work without it
    finally:
        eng.close()
