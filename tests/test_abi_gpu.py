"""C-ABI robustness on a device: descriptors handed to sdm_load_weights are validated before anything is dereferenced,
a failed (re)load leaves the handle "not loaded" (never half-loaded), options are checked."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _desc_array(E, items):
    arr = (E.sdm_tensor_desc * len(items))()
    keep = []
    for i, (name, dtype, ndim, shape, t) in enumerate(items):
        arr[i].name = name
        arr[i].dtype = dtype
        arr[i].ndim = ndim
        for j, s in enumerate(shape):
            arr[i].shape[j] = s
        arr[i].data = t.data_ptr() if t is not None else None
        keep.append(t)
    return arr, keep


@pytest.mark.parametrize("case", ["rank5", "bad_dtype", "null_name", "null_data", "negative_extent"])
def test_load_weights_validates_descriptors(pkg, case):
    E = pkg.engine
    lib = E.load_library()
    h = C.c_void_p()
    assert lib.sdm_create(C.byref(h), 0) == 0
    t = torch.zeros(4, 4)
    item = {
        "rank5": (b"unet.x", 0, 5, (1, 1, 1, 1), t),
        "bad_dtype": (b"unet.x", 7, 2, (4, 4), t),
        "null_name": (None, 0, 2, (4, 4), t),
        "null_data": (b"unet.x", 0, 2, (4, 4), None),
        "negative_extent": (b"unet.x", 0, 2, (-4, 4), t),
    }[case]
    arr, keep = _desc_array(E, [item])
    try:
        rc = lib.sdm_load_weights(h, arr, 1)
        assert rc != 0
        msg = lib.sdm_last_error().decode()
        assert "sdm_load_weights" in msg, msg
    finally:
        lib.sdm_destroy(h)
    del keep


def test_failed_reload_leaves_the_engine_unloaded(pkg):
    """engine_load clears `loaded` first: after a reload that fails (here: a checkpoint with keys missing) the handle reports
    'weights not loaded' instead of running with freed buffers."""
    from oracle import synth

    ckpt = synth.make_checkpoint(seed=1234)
    eng = pkg.engine.Engine(0)
    eng.load_state_dict(ckpt)
    image, trimap = synth.make_inputs(1, 64, seed=1)
    eng.forward(image.cuda(), trimap.cuda(), False)
    broken = {k: v for k, v in ckpt.items() if not k.startswith("vae.decoder.conv_out")}
    with pytest.raises(RuntimeError, match="missing"):
        eng.load_state_dict(broken)
    with pytest.raises(RuntimeError, match="not loaded"):
        eng.forward(image.cuda(), trimap.cuda(), False)
    eng.load_state_dict(ckpt)  # and it recovers
    eng.forward(image.cuda(), trimap.cuda(), False)
    with pytest.raises(RuntimeError, match="unknown engine option"):
        eng.set_option("no_such_option", 1)
    eng.close()
