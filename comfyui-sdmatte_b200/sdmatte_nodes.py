"""`Apply SDMatte` ComfyUI node backed by the B200-native engine.

Drop-in for the node surface of the reference (/root/reference/sdmatte_nodes.py:217-414): same class name, widget
schema, RETURN_TYPES / RETURN_NAMES / FUNCTION / CATEGORY, same `apply_matte` signature and outputs.  What changes is
what sits behind it: instead of building a diffusers model and calling it under autocast (sdmatte_nodes.py:286-360), the
loader hands the checkpoint to `engine.Engine` (hand-written sm_100a kernels behind a C ABI) and caches it per
(checkpoint, device).  There is no CPU path: `force_cpu=True` raises.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch

from . import engine as _engine

try:  # inside ComfyUI
    import folder_paths  # type: ignore
    import comfy.model_management as _mm  # type: ignore
except Exception:  # outside ComfyUI (tests, bench): minimal stand-ins with the same calls the node makes
    class _FolderPaths:
        models_dir = os.environ.get("SDMATTE_MODELS_DIR", os.path.join(os.path.dirname(os.path.abspath(__file__)), "models"))
        _paths: Dict[str, list] = {}

        def add_model_folder_path(self, name, path):
            self._paths.setdefault(name, [])
            if path not in self._paths[name]:
                self._paths[name].append(path)

        def get_folder_paths(self, name):
            return list(self._paths.get(name, []))

    class _MM:
        @staticmethod
        def get_torch_device():
            return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    folder_paths = _FolderPaths()  # type: ignore
    _mm = _MM()  # type: ignore

MODEL_DIR = os.path.join(folder_paths.models_dir, "SDMatte")
folder_paths.add_model_folder_path("SDMatte", MODEL_DIR)

# same checkpoint names as the reference's MODEL_URLS (sdmatte_nodes.py:14-17); downloading is out of scope here
CKPT_NAMES = ["SDMatte.safetensors", "SDMatte_plus.safetensors"]

# tests / bench can register an in-memory state dict under a checkpoint name (no file needed)
_STATE_DICT_OVERRIDES: Dict[str, Dict[str, torch.Tensor]] = {}
_ENGINE_CACHE: Dict[Tuple[str, str], "_engine.Engine"] = {}


def register_state_dict(ckpt_name: str, state_dict: Dict[str, torch.Tensor]) -> None:
    _STATE_DICT_OVERRIDES[ckpt_name] = state_dict
    for k in [k for k in _ENGINE_CACHE if k[0] == ckpt_name]:
        _ENGINE_CACHE.pop(k).close()


def find_checkpoint(model_name: str) -> str:
    """Search order of the reference's download_model (sdmatte_nodes.py:103-130): registered SDMatte folders, then MODEL_DIR."""
    for search_path in (folder_paths.get_folder_paths("SDMatte") or []) + [MODEL_DIR]:
        p = os.path.join(search_path, model_name)
        try:
            if os.path.isfile(p) and os.path.getsize(p) > 0:
                return p
        except OSError:
            pass
    if model_name not in CKPT_NAMES:
        raise ValueError(f"[SDMatte] Unknown model name: {model_name}")
    raise FileNotFoundError(f"[SDMatte] '{model_name}' not found under {folder_paths.get_folder_paths('SDMatte')}; "
                            "place the checkpoint there (this build does not download)")


def _load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    from safetensors import safe_open

    sd = {}
    with safe_open(path, framework="pt", device="cpu") as f:
        for key in f.keys():
            sd[key] = f.get_tensor(key)
    return sd


def get_engine(ckpt_name: str, device: torch.device) -> "_engine.Engine":
    key = (ckpt_name, str(device))
    eng = _ENGINE_CACHE.get(key)
    if eng is None:
        sd = _STATE_DICT_OVERRIDES.get(ckpt_name)
        eng = _engine.Engine(device)
        if sd is not None:
            used, unexpected = eng.load_state_dict(sd)
        elif os.environ.get("SDMATTE_PY_SAFETENSORS") == "1":  # the safetensors package instead of the native reader
            used, unexpected = eng.load_state_dict(_load_state_dict(find_checkpoint(ckpt_name)))
        else:  # native reader: header parse + mmap, repacked straight from the mapping (no torch CPU tensors)
            used, unexpected = eng.load_safetensors(find_checkpoint(ckpt_name))
        print(f"[SDMatte-B200] loaded {ckpt_name}: {used} tensors used, {unexpected} ignored")
        _ENGINE_CACHE[key] = eng
    return eng


class SDMatteApply:
    @classmethod
    def INPUT_TYPES(s):
        return {
            "required": {
                "ckpt_name": (list(CKPT_NAMES),),
                "image": ("IMAGE", {"tooltip": "input image to matte"}),
                "trimap": ("MASK", {"tooltip": "trimap mask: white = foreground, black = background, gray = unknown"}),
                "inference_size": ([512, 640, 768, 896, 1024], {"default": 1024, "tooltip": "inference resolution"}),
                "is_transparent": ("BOOLEAN", {"default": False, "tooltip": "input image contains a transparent object"}),
                "output_mode": (["alpha_only", "matted_rgba", "matted_rgb"], {"default": "alpha_only"}),
                "mask_refine": ("BOOLEAN", {"default": True, "tooltip": "filter the matte with the trimap"}),
                "trimap_constraint": ("FLOAT", {"default": 0.8, "min": 0.1, "max": 1.0, "step": 0.1}),
            },
            "optional": {
                "force_cpu": ("BOOLEAN", {"default": False}),
            },
        }

    RETURN_TYPES = ("MASK", "IMAGE")
    RETURN_NAMES = ("alpha_mask", "matted_image")
    FUNCTION = "apply_matte"
    CATEGORY = "Matting/SDMatte"

    def apply_matte(self, ckpt_name, image, trimap, inference_size, is_transparent, output_mode, mask_refine, trimap_constraint,
                    force_cpu=False):
        if force_cpu:
            raise RuntimeError("[SDMatte-B200] force_cpu is not supported: this node runs only on an sm_100a (B200) GPU")
        device = _mm.get_torch_device()
        if torch.device(device).type != "cuda":
            raise RuntimeError("[SDMatte-B200] no CUDA device available; there is no CPU fallback")
        device = torch.device(device)
        if image.dim() != 4 or image.shape[-1] != 3:
            raise ValueError(f"[SDMatte] image must be (B,H,W,3), got {tuple(image.shape)}")
        if trimap.dim() != 3 or trimap.shape[0] != image.shape[0] or trimap.shape[1:] != image.shape[1:3]:
            raise ValueError(f"[SDMatte] trimap must be (B,H,W) matching the image, got {tuple(trimap.shape)} vs {tuple(image.shape)}")
        eng = get_engine(ckpt_name, device)
        B, H, W, _ = image.shape
        R = int(inference_size)

        # pre-processing (sdmatte_nodes.py:339-353): one H2D of the caller's tensors, antialiased-bilinear resize to R x R on
        # the device (csrc/prepost.cu); the (x-0.5)/0.5 and *2-1 normalisations happen inside the engine
        img = image.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        tri = trimap.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        img_r, tri_r = _engine.preprocess(img, tri, R)
        alpha = eng.forward(img_r, tri_r, bool(is_transparent))  # (B,R,R) fp16, in [0,1]

        # post-processing (sdmatte_nodes.py:362-397) in ONE kernel: resize back, clamp, mask_refine, composition; then one
        # D2H per output.  fp16 alpha like the reference's CUDA path.
        out_d, matted_d = _engine.postprocess(alpha, img, tri, output_mode, bool(mask_refine), float(trimap_constraint))
        out = out_d.cpu()
        if matted_d is None:  # "alpha_only": zeros_like(image) (sdmatte_nodes.py:384-385)
            matted = torch.zeros_like(image, device="cpu")
        else:
            matted = matted_d.cpu()
        return (out, matted)


NODE_CLASS_MAPPINGS = {"SDMatteApply": SDMatteApply}
NODE_DISPLAY_NAME_MAPPINGS = {"SDMatteApply": "Apply SDMatte"}
