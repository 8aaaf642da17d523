"""TEST INFRASTRUCTURE: execute the reference's OWN functions from /root/reference without importing diffusers/comfy.

The reference modules cannot be imported here (diffusers, comfy, folder_paths are absent), but the functions on the matte
path that do not touch diffusers internals can be lifted by parsing the source with `ast` and exec'ing just those
function definitions.  Nothing is copied into this repository: the source is read from /root/reference at run time, so
this only works in the build container (the GPU box has no /root/reference; goldens produced here are committed under
tests/golden/ by tests/golden/make_golden.py).
"""
from __future__ import annotations

import ast
import math
import os
from types import SimpleNamespace

import torch
import torch.nn.functional as F

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "utils"))


def _lift(path: str, names, extra_globals=None):
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F, "math": math, "nn": torch.nn}
    if extra_globals:
        ns.update(extra_globals)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
    return SimpleNamespace(**{n: ns[n] for n in names})


def attention_fns():
    """custom_prepare_attention_mask, custom_get_attention_scores  (src/utils/replace.py:20-122)"""
    return _lift(os.path.join(REF, "src/utils/replace.py"), ["custom_prepare_attention_mask", "custom_get_attention_scores"])


def surgery_fns():
    """replace_unet_conv_in, add_aux_conv_in  (src/utils/utils.py:13-41)"""
    from torch.nn import Conv2d
    from torch.nn.parameter import Parameter

    return _lift(os.path.join(REF, "src/utils/utils.py"), ["replace_unet_conv_in", "add_aux_conv_in"],
                 {"Conv2d": Conv2d, "Parameter": Parameter})


def node_helpers():
    """_resize_norm_image_bchw, _resize_mask_b1hw  (sdmatte_nodes.py:204-214)"""
    from torchvision import transforms

    return _lift(os.path.join(REF, "sdmatte_nodes.py"), ["_resize_norm_image_bchw", "_resize_mask_b1hw"], {"transforms": transforms})


def node_postprocess(pred_alpha, image, trimap, output_mode, mask_refine, trimap_constraint):
    """Run the reference's post-processing statements (sdmatte_nodes.py:362-397) verbatim on the given tensors by
    extracting them from apply_matte's body."""
    from torchvision import transforms

    path = os.path.join(REF, "sdmatte_nodes.py")
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    body = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "apply_matte":
            body = node.body
    assert body is not None
    # statements from `out = transforms.Resize((orig_h, orig_w))(pred_alpha)` up to (not including) the final cuda cleanup
    start = next(i for i, st in enumerate(body) if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "out")
    end = next(i for i, st in enumerate(body) if i > start and isinstance(st, ast.If) and "device" in ast.unparse(st.test) and i > start + 3)
    mod = ast.Module(body=body[start:end], type_ignores=[])
    ns = {"torch": torch, "transforms": transforms, "pred_alpha": pred_alpha, "image": image, "trimap": trimap,
          "orig_h": image.shape[1], "orig_w": image.shape[2], "output_mode": output_mode, "mask_refine": mask_refine,
          "trimap_constraint": trimap_constraint}
    exec(compile(mod, path, "exec"), ns)
    return ns["out"], ns["matted_image"]
