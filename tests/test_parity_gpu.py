"""Parity at the BASELINE resolutions (R = 512 / 1024, bs = 1 / 8) against an INDEPENDENT fp16 reference on the same B200.

The checker is the oracle graph (oracle/sdmatte_oracle.py, follows meta_arch.py:127-261 / replace.py:379-549) executed ON THE GPU
in two more modes — cuDNN / cuBLAS are allowed there, it is test infrastructure:
  * fp32      : everything fp32, TF32 off (what the reference's CPU branch computes, sdmatte_nodes.py:359-360);
  * autocast  : genuine `torch.autocast("cuda", torch.float16)` over fp32 master weights with SlicedAttnProcessor(1) semantics —
                the reference's real CUDA branch, sdmatte_nodes.py:331-358.  Its rounding points are PyTorch's, not ours.
north_star asks |d_alpha| <= 1e-3 against the reference.  What can be checked here: the reference's own CUDA branch (autocast)
sits `floor = autocast<->fp32` away from exact arithmetic on this checkpoint; an independent fp16 implementation of the same graph
cannot be closer to autocast than two such floors, nor meaningfully closer to fp32 than one.  The asserts therefore are
expressed against the MEASURED floor at the same (R, B): engine<->fp32 mean and 99.9th percentile <= 1.25 x floor, max <= 1.5 x
floor (the maximum over 1e6..8e6 pixels of two independent noise realisations is itself noisy), and every figure is recorded in
gpurun_out/parity_r2.json -> profiles/ for DESIGN.md.  Per-block error growth: engine taps (sdm_set_option keep_taps) against
the same two checkers, relative RMS error per block, recorded as a curve and bounded by 1.5 x the autocast curve.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "parity_r2.json")


def _record(name, payload):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    data = {}
    if os.path.exists(OUT):
        try:
            data = json.load(open(OUT))
        except Exception:
            data = {}
    data[name] = payload
    json.dump(data, open(OUT, "w"), indent=1)


@pytest.fixture(scope="module")
def ckpt():
    from oracle import synth

    return synth.make_checkpoint(seed=1234)


@pytest.fixture(scope="module")
def sd_gpu(ckpt):
    from oracle import sdmatte_oracle as orc

    return orc.to_device(ckpt, "cuda")


@pytest.fixture(scope="module")
def engine(pkg, ckpt):
    eng = pkg.engine.Engine(0)
    eng.load_state_dict(ckpt)
    yield eng
    eng.close()


def _stats(a, b):
    d = (a.float() - b.float()).abs().flatten()
    k = max(1, int(d.numel() * 0.999))
    return {"max": d.max().item(), "p999": d.kthvalue(k).values.item(), "mean": d.mean().item()}


def _oracle_per_sample(sd_gpu, image, trimap, flags, mode, capture=None):
    """The checker one sample at a time (samples are independent; keeps the fp32 activations of a 1024^2 batch off the device)."""
    from oracle import sdmatte_oracle as orc

    alphas, means = [], []
    for b in range(image.shape[0]):
        r = orc.forward(sd_gpu, image[b:b + 1], trimap[b:b + 1], is_transparent=[flags[b]], mode=mode, sliced=True,
                        capture=capture if b == 0 else None, device="cuda")
        alphas.append(r["alpha"].float().squeeze(1))
        means.append(r["label_mean"].float().squeeze(1))
        del r
    torch.cuda.empty_cache()
    return torch.cat(alphas), torch.cat(means)


@pytest.mark.parametrize("R,B", [(512, 1), (1024, 1), (1024, 8)])  # BASELINE.json configs 2 / 3 (+ the 512^2 case of config 1)
def test_alpha_vs_gpu_checkers_at_baseline_sizes(engine, sd_gpu, R, B):
    from oracle import synth

    image, trimap = synth.make_inputs(B, R, seed=R + B)
    flags = [False] * B  # BASELINE configs 2 / 3: is_transparent=False
    alpha, pre = engine.forward(image.cuda(), trimap.cuda(), flags, want_premean=True)
    alpha, pre = alpha.float(), pre.float()
    a32, m32 = _oracle_per_sample(sd_gpu, image, trimap, flags, "fp32")
    a16, m16 = _oracle_per_sample(sd_gpu, image, trimap, flags, "autocast")
    row = {
        "engine_vs_fp32": _stats(alpha, a32), "engine_vs_autocast": _stats(alpha, a16), "autocast_vs_fp32": _stats(a16, a32),
        "premean_engine_vs_fp32": _stats(pre, m32), "premean_engine_vs_autocast": _stats(pre, m16), "premean_autocast_vs_fp32": _stats(m16, m32),
        "alpha_std": a32.std().item(), "alpha_saturated_frac": ((a32 == 0) | (a32 == 1)).float().mean().item(),
    }
    for k, v in row.items():
        print(f"[parity R={R} B={B}] {k}: {v}")
    _record(f"alpha_R{R}_B{B}", row)
    floor, eng = row["autocast_vs_fp32"], row["engine_vs_fp32"]
    assert eng["mean"] <= 1.25 * floor["mean"], (eng, floor)
    assert eng["p999"] <= 1.25 * floor["p999"], (eng, floor)
    assert eng["max"] <= 1.5 * floor["max"], (eng, floor)
    # and against the fp16 reference itself: two independent fp16 realisations are at most two floors apart
    ea = row["engine_vs_autocast"]
    assert ea["mean"] <= 2.0 * floor["mean"] and ea["max"] <= 2.0 * floor["max"], (ea, floor)
    pm, pf = row["premean_engine_vs_fp32"], row["premean_autocast_vs_fp32"]
    assert pm["mean"] <= 1.25 * pf["mean"] and pm["max"] <= 1.5 * pf["max"], (pm, pf)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).float()


def _rel_rms(got, want):
    return ((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt().clamp_min(1e-20)).item()


@pytest.mark.parametrize("R", [512, 1024])
def test_per_block_error_growth(pkg, ckpt, sd_gpu, R):
    """One tap per block (VAE levels, every UNet block, decoder levels): relative RMS error of the engine and of the autocast
    checker against fp32, in graph order."""
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    B = 1
    image, trimap = synth.make_inputs(B, R, seed=R + 7)
    eng = pkg.engine.Engine(0)
    eng.load_state_dict(ckpt)
    eng.set_option("keep_taps", 1)
    try:
        alpha = eng.forward(image.cuda(), trimap.cuda(), False).float()
        names = eng.tap_names()
        cap32, cap16 = {}, {}
        r32 = orc.forward(sd_gpu, image, trimap, mode="fp32", sliced=True, capture=cap32, device="cuda")
        r16 = orc.forward(sd_gpu, image, trimap, mode="autocast", capture=cap16, device="cuda")
        curve = []
        for n in names:
            got = eng.debug_tensor(n).float()
            if n.startswith("enc."):  # engine batch = [rgb samples ; trimap samples]
                key = n[4:]
                w32 = torch.cat([_nhwc(cap32["enc_rgb." + key]), _nhwc(cap32["enc_tri." + key])])
                w16 = torch.cat([_nhwc(cap16["enc_rgb." + key]), _nhwc(cap16["enc_tri." + key])])
            elif n == "ctx":
                w32, w16 = cap32[n].float().reshape(got.shape), cap16[n].float().reshape(got.shape)
            else:
                w32, w16 = _nhwc(cap32[n]), _nhwc(cap16[n])
            assert got.shape == w32.shape, (n, got.shape, w32.shape)
            e = {"tap": n, "engine_vs_fp32": _rel_rms(got, w32), "autocast_vs_fp32": _rel_rms(w16, w32), "engine_vs_autocast": _rel_rms(got, w16),
                 "engine_max_rel_to_range": ((got - w32).abs().max() / w32.abs().max()).item()}
            curve.append(e)
            print(f"[growth R={R}] {n:18s} engine {e['engine_vs_fp32']:.3e}  autocast {e['autocast_vs_fp32']:.3e}  eng-vs-ac {e['engine_vs_autocast']:.3e}  max/range {e['engine_max_rel_to_range']:.3e}")
            del got, w32, w16
        fin = {"engine_vs_fp32": _stats(alpha, r32["alpha"].float().squeeze(1)), "autocast_vs_fp32": _stats(r16["alpha"].float().squeeze(1), r32["alpha"].float().squeeze(1))}
        _record(f"error_growth_R{R}", {"curve": curve, "alpha": fin})
        assert len(curve) >= 40
        for e in curve:
            assert e["engine_vs_fp32"] <= 1.5 * e["autocast_vs_fp32"] + 1e-4, e
    finally:
        eng.close()


def test_fp16sim_emulation_vs_genuine_autocast(sd_gpu):
    """How good is the hand-written rounding-point emulation (mode fp16sim) that round 1 used as its fp16 floor?  Against the genuine
    autocast run at R=512: recorded, and bounded by two floors."""
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    R = 512
    image, trimap = synth.make_inputs(1, R, seed=5)
    a32 = orc.forward(sd_gpu, image, trimap, mode="fp32", sliced=True, device="cuda")["alpha"].float()
    a16 = orc.forward(sd_gpu, image, trimap, mode="autocast", device="cuda")["alpha"].float()
    asim = orc.forward(sd_gpu, image, trimap, mode="fp16sim", sliced=True, device="cuda")["alpha"].float()
    row = {"autocast_vs_fp32": _stats(a16, a32), "fp16sim_vs_fp32": _stats(asim, a32), "autocast_vs_fp16sim": _stats(a16, asim)}
    print("[fp16sim]", row)
    _record("fp16sim_R512", row)
    assert row["fp16sim_vs_fp32"]["mean"] <= 1.5 * row["autocast_vs_fp32"]["mean"]
    assert row["autocast_vs_fp16sim"]["mean"] <= 2.0 * row["autocast_vs_fp32"]["mean"]


def test_gpu_fp32_checker_equals_cpu_oracle(ckpt, sd_gpu):
    """The GPU-resident fp32 checker is the same restatement as the CPU oracle the golden vectors pin (tests/test_oracle_cpu.py):
    same graph, TF32 off; they differ by summation order only."""
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    image, trimap = synth.make_inputs(1, 64, seed=0)
    cpu = orc.forward(ckpt, image, trimap)
    gpu = orc.forward(sd_gpu, image, trimap, device="cuda")
    d = (cpu["alpha"] - gpu["alpha"].cpu()).abs().max().item()
    dm = (cpu["label_mean"] - gpu["label_mean"].cpu()).abs().max().item()
    print(f"[checker cpu vs gpu fp32] max|da|={d:.3e} max|dmean|={dm:.3e}")
    assert d <= 2e-5 and dm <= 5e-5


def test_no_foreground_key_fp16_score_quantisation(engine, sd_gpu):
    """with NO foreground key in a sample every key carries a -5000 / -10000 bias and the reference's autocast branch
    rounds (scale q.k + bias) to fp16, whose spacing is 4 / 8 there: the scores collapse onto a few values and the softmax
    becomes near-uniform over the max-bias keys, while fp32 arithmetic (the reference's CPU branch, and the engine) keeps the
    q.k ordering.  Quantifies the deviation; the engine is asserted against the fp32 branch."""
    from oracle import synth

    R = 256
    image, trimap = synth.make_inputs(1, R, seed=13)
    trimap = trimap.clamp(max=0.5)  # no pure-white key anywhere
    alpha = engine.forward(image.cuda(), trimap.cuda(), False).float()
    a32, _ = _oracle_per_sample(sd_gpu, image, trimap, [False], "fp32")
    a16, _ = _oracle_per_sample(sd_gpu, image, trimap, [False], "autocast")
    row = {"engine_vs_fp32": _stats(alpha, a32), "engine_vs_autocast": _stats(alpha, a16), "autocast_vs_fp32": _stats(a16, a32)}
    print("[no-fg]", row)
    _record("no_foreground_R256", row)
    assert row["engine_vs_fp32"]["max"] <= 6e-3 and row["engine_vs_fp32"]["mean"] <= 6e-4
