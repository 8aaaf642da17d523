"""`Apply SDMatte` ComfyUI node backed by the B200-native engine.

Drop-in for the node surface of the reference (/root/reference/sdmatte_nodes.py:217-414): same class name, widget
schema, RETURN_TYPES / RETURN_NAMES / FUNCTION / CATEGORY, same `apply_matte` signature and outputs.  What changes is
what sits behind it: instead of building a diffusers model and calling it under autocast (sdmatte_nodes.py:286-360), the
loader hands the checkpoint to `engine.Engine` (hand-written sm_100a kernels behind a C ABI) and caches it per
(checkpoint, device).  There is no CPU path: `force_cpu=True` raises.

Data path of one call (everything between the argument checks and the return is ONE C-ABI call per device,
`sdm_apply_matte_host`): the caller's pageable host tensors are staged through a page-locked buffer by several host threads
with the H2D copies enqueued chunk by chunk -> resize -> forward (CUDA-graph replay from the second call of a geometry on) ->
post-processing -> D2H.  With several devices configured (`set_devices` / SDMATTE_DEVICES) a batch of B >= 2 samples is
sharded contiguously over min(B, N) GPUs, one host thread + one engine handle per GPU inside this process (SURVEY §8(b),(e));
samples are independent, so the sharded result is bit-identical to the single-GPU one (tests/test_multigpu_gpu.py).
"""
from __future__ import annotations

import os
import sys
import threading
import time
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import assets as _assets
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

# same checkpoint names / URLs as the reference's MODEL_URLS (sdmatte_nodes.py:14-17)
MODEL_URLS = _assets.MODEL_URLS
CKPT_NAMES = list(MODEL_URLS)

# tests / bench can register an in-memory state dict under a checkpoint name (no file needed)
_STATE_DICT_OVERRIDES: Dict[str, Dict[str, torch.Tensor]] = {}
_ENGINE_CACHE: Dict[Tuple[str, str], "_engine.Engine"] = {}
_CACHE_LOCK = threading.Lock()
_DEVICES: Optional[List[torch.device]] = None


def register_state_dict(ckpt_name: str, state_dict: Dict[str, torch.Tensor]) -> None:
    _STATE_DICT_OVERRIDES[ckpt_name] = state_dict
    with _CACHE_LOCK:
        for k in [k for k in _ENGINE_CACHE if k[0] == ckpt_name]:
            _ENGINE_CACHE.pop(k).close()


def unload_engines() -> None:
    """Drop every cached engine: weights (1.9 GB per GPU), workspace and page-locked staging.  The reference frees its model after
    every call (sdmatte_nodes.py:399-403); this node keeps it resident for the next call unless SDMATTE_KEEP_RESIDENT=0 or the
    host calls this hook (e.g. from ComfyUI's "unload models")."""
    with _CACHE_LOCK:
        for k in list(_ENGINE_CACHE):
            eng = _ENGINE_CACHE.pop(k)
            eng._ws = None
            eng.close()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def set_devices(devices: Optional[Sequence]) -> None:
    """GPUs a batch may be sharded over (None: only the device ComfyUI hands out).  Also settable with SDMATTE_DEVICES="0,1,2,3"
    or "all"."""
    global _DEVICES
    _DEVICES = None if devices is None else [torch.device("cuda", int(d)) if not isinstance(d, torch.device) else d for d in devices]


def _configured_devices(primary: torch.device) -> List[torch.device]:
    if _DEVICES is not None:
        return list(_DEVICES)
    env = os.environ.get("SDMATTE_DEVICES", "").strip()
    if env == "all":
        return [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
    if env:
        return [torch.device("cuda", int(x)) for x in env.split(",") if x.strip() != ""]
    return [primary]


def shard_bounds(B: int, n: int) -> List[Tuple[int, int]]:
    """Contiguous batch shards for n devices (sizes differ by at most one; empty shards are dropped)."""
    n = max(1, min(n, B))
    base, extra = divmod(B, n)
    out, lo = [], 0
    for r in range(n):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def find_checkpoint(model_name: str) -> str:
    """download_model of the reference (sdmatte_nodes.py:103-199): registered SDMatte folders, then MODEL_DIR, then download."""
    return _assets.download_model(model_name, MODEL_DIR, folder_paths.get_folder_paths("SDMatte") or [])


def _load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    from safetensors import safe_open

    sd = {}
    with safe_open(path, framework="pt", device="cpu") as f:
        for key in f.keys():
            sd[key] = f.get_tensor(key)
    return sd


def get_engine(ckpt_name: str, device: torch.device) -> "_engine.Engine":
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (ckpt_name, str(device))
    with _CACHE_LOCK:
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
        print(f"[SDMatte-B200] loaded {ckpt_name} on {device}: {used} tensors used, {unexpected} ignored", file=sys.stderr)
        with _CACHE_LOCK:
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
        B, H, W, _ = image.shape
        R = int(inference_size)
        # inputs are borrowed and never written; host fp32 contiguous is what ComfyUI hands over (no copy in that case)
        img = image.detach().to(device="cpu", dtype=torch.float32).contiguous()
        tri = trimap.detach().to(device="cpu", dtype=torch.float32).contiguous()
        mode = _engine.OUTPUT_MODES.get(output_mode, 3)
        alpha = torch.empty((B, H, W), dtype=torch.float16)  # fp16 like the reference's CUDA branch
        zeros_box: List[torch.Tensor] = []
        zeros_thread = None
        if mode == 0:
            # "alpha_only": zeros_like(image) (sdmatte_nodes.py:384-385).  Clearing 100 MB of host memory takes ~10 ms of one
            # core: it runs on a helper thread while the GPU works (the library call below releases the GIL), and it starts
            # after the input staging, which needs the host memory bandwidth for itself (r2j: 7.2 vs 4.4 ms of staging)
            def _zeros():
                if img.numel() * 4 > (32 << 20):
                    time.sleep(0.015)
                zeros_box.append(torch.zeros_like(img))

            zeros_thread = threading.Thread(target=_zeros)
            zeros_thread.start()
            matted = None
        else:
            matted = torch.empty((B, H, W, 4 if mode == 1 else 3), dtype=torch.float32)

        devices = _configured_devices(device)
        shards = shard_bounds(B, len(devices))

        def run(dev, lo, hi):
            eng = get_engine(ckpt_name, dev)
            eng.apply_host(img[lo:hi], tri[lo:hi], R, bool(is_transparent), output_mode, bool(mask_refine), float(trimap_constraint),
                           alpha_out=alpha[lo:hi], matted_out=None if mode == 0 else matted[lo:hi])

        if len(shards) == 1:
            run(devices[0], 0, B)
        else:  # one host thread per GPU; the library call releases the GIL
            errors: List[BaseException] = []

            def guarded(dev, lo, hi):
                try:
                    run(dev, lo, hi)
                except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                    errors.append(e)

            threads = [threading.Thread(target=guarded, args=(devices[r], lo, hi)) for r, (lo, hi) in enumerate(shards)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
        if zeros_thread is not None:
            zeros_thread.join()
            matted = zeros_box[0]
        if os.environ.get("SDMATTE_KEEP_RESIDENT", "1") == "0":
            unload_engines()
        return (alpha, matted)


NODE_CLASS_MAPPINGS = {"SDMatteApply": SDMatteApply}
NODE_DISPLAY_NAME_MAPPINGS = {"SDMatteApply": "Apply SDMatte"}
