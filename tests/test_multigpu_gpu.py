"""Multi-GPU behind the node (reference contract: the per-sample independence of sdmatte_nodes.py:339-363):
one process, one engine handle + one host thread per GPU, the batch sharded contiguously.  The N-way sharded result must be
BIT-IDENTICAL to the single-GPU result (samples are independent and the kernel choice depends on per-sample geometry only).
Skipped on boxes with fewer than two GPUs (`gpurun --gpus 2`).  The node-call tests at the bottom need one GPU only.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ckpt():
    from oracle import synth

    return synth.make_checkpoint(seed=1234)


@pytest.fixture()
def nodes(pkg, ckpt):
    n = pkg.sdmatte_nodes
    n.register_state_dict("SDMatte.safetensors", ckpt)
    yield n
    n.set_devices(None)
    n.unload_engines()


def test_sharded_batch_is_bit_identical_to_single_gpu(nodes):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import synth

    node = nodes.SDMatteApply()
    B, R = 2 * ngpu + 1, 256  # uneven split on purpose: shard sizes differ by one
    image, trimap = synth.make_inputs(B, R, seed=77, Hin=300, Win=200)
    nodes.set_devices([0])
    a1, m1 = node.apply_matte("SDMatte.safetensors", image, trimap, R, False, "matted_rgba", True, 0.8)
    nodes.set_devices(list(range(ngpu)))
    aN, mN = node.apply_matte("SDMatte.safetensors", image, trimap, R, False, "matted_rgba", True, 0.8)
    assert torch.equal(a1, aN), "sharded alpha differs from the single-GPU alpha"
    assert torch.equal(m1, mN)
    # a second device alone (kernel function attributes are per device) gives the same bits as device 0
    nodes.set_devices([1])
    a2, _ = node.apply_matte("SDMatte.safetensors", image[:2], trimap[:2], R, False, "alpha_only", True, 0.8)
    assert torch.equal(a2, a1[:2])


def test_shard_bounds_cover_the_batch(pkg):
    sb = pkg.sdmatte_nodes.shard_bounds
    for B in (1, 2, 7, 8, 64):
        for n in (1, 2, 3, 4, 8):
            s = sb(B, n)
            assert s[0][0] == 0 and s[-1][1] == B and all(a[1] == b[0] for a, b in zip(s, s[1:]))
            assert all(hi > lo for lo, hi in s) and max(hi - lo for lo, hi in s) - min(hi - lo for lo, hi in s) <= 1


def test_node_call_equals_stepwise_pipeline(pkg, nodes):
    """sdm_apply_matte_host (the node's single library call: staging, H2D, resize, forward, post-processing, D2H) == the same
    stages called one by one through sdm_preprocess / sdm_forward / sdm_postprocess, bit for bit; pageable and pinned inputs agree;
    the second call of a geometry replays the CUDA graph and still agrees."""
    from oracle import synth

    eng_mod = pkg.engine
    R = 128
    image, trimap = synth.make_inputs(2, R, seed=23, Hin=150, Win=100)
    eng = nodes.get_engine("SDMatte.safetensors", torch.device("cuda", 0))
    for mode in ("alpha_only", "matted_rgba", "matted_rgb"):
        a_host, m_host = eng.apply_host(image, trimap, R, [False, True], mode, True, 0.8)
        img_d, tri_d = image.cuda(), trimap.cuda()
        img_r, tri_r = eng_mod.preprocess(img_d, tri_d, R)
        alpha = eng.forward(img_r, tri_r, [False, True])
        a_dev, m_dev = eng_mod.postprocess(alpha, img_d, tri_d, mode, True, 0.8)
        assert torch.equal(a_host, a_dev.cpu()), mode
        if mode != "alpha_only":
            assert torch.equal(m_host, m_dev.cpu()), mode
        else:
            assert m_host is None
    g0 = eng.graph_stats()
    a1, _ = eng.apply_host(image, trimap, R, False, "alpha_only", True, 0.8)
    a2, _ = eng.apply_host(image.pin_memory(), trimap.pin_memory(), R, False, "alpha_only", True, 0.8)
    a3, _ = eng.apply_host(image, trimap, R, False, "alpha_only", True, 0.8)
    g1 = eng.graph_stats()
    assert torch.equal(a1, a2) and torch.equal(a1, a3)
    assert g1["launches"] > g0["launches"], "repeated node calls of one geometry should replay the plan's CUDA graph"
    # inputs already at R x R: no resize kernels on the path, same call
    image2, trimap2 = synth.make_inputs(1, R, seed=24)
    a4, _ = eng.apply_host(image2, trimap2, R, False, "alpha_only", False, 0.8)
    a5 = eng.forward(image2.cuda(), trimap2.cuda(), False)
    assert torch.equal(a4, a5.cpu())


def test_cuda_graph_replay_equals_eager(pkg, ckpt):
    from oracle import synth

    image, trimap = synth.make_inputs(2, 128, seed=41)
    img, tri = image.cuda(), trimap.cuda()
    eager = pkg.engine.Engine(0)
    eager.load_state_dict(ckpt)
    eager.set_option("cuda_graph", 0)
    a_ref = eager.forward(img, tri, [True, False]).clone()
    assert eager.graph_stats()["captures"] == 0
    eager.close()
    eng = pkg.engine.Engine(0)
    eng.load_state_dict(ckpt)
    out = torch.empty((2, 128, 128), dtype=torch.float16, device="cuda")
    for i in range(4):  # 1st eager, 2nd captures + launches, 3rd / 4th replay
        out.zero_()
        eng.forward(img, tri, [True, False], out=out)
        assert torch.equal(out, a_ref), f"call {i}"
    g = eng.graph_stats()
    assert g["captures"] == 1 and g["launches"] == 3, g
    # changed flags: same graph (the flags live in device memory), new values
    eng.forward(img, tri, [False, False], out=out)
    a_ff = eager_like = out.clone()
    assert not torch.equal(a_ff, a_ref)
    assert eng.graph_stats()["captures"] == 1
    eng.close()
    del eager_like
