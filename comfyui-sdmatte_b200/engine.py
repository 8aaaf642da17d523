"""ctypes binding of libsdmatte_b200.so (include/sdmatte_b200.h) + a thin torch-tensor convenience layer.

PyTorch is used for device memory and streams only; every computation happens inside the C-ABI library.
There is no CPU fallback: if the CUDA extension is missing or the device is not sm_100, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsdmatte_b200.so")
_lib = None
ABI_VERSION = 201  # sdm_version(): bumped whenever a struct or signature of include/sdmatte_b200.h changes


class sdm_tensor_desc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int), ("ndim", C.c_int), ("shape", C.c_int64 * 4), ("data", C.c_void_p)]


class sdm_conv_gemm_args(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("nsrc", C.c_int),
        ("src0", C.c_void_p), ("c0", C.c_int), ("ld0", C.c_int64),
        ("src1", C.c_void_p), ("c1", C.c_int), ("ld1", C.c_int64),
        ("ksize", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("w", C.c_void_p), ("N", C.c_int), ("w_bstride", C.c_int64),
        ("mode", C.c_int), ("ups2", C.c_int),
        ("out", C.c_void_p), ("out_ld", C.c_int64), ("out_bstride", C.c_int64),
        ("bias", C.c_void_p), ("bias_sel", C.c_void_p),
        ("res", C.c_void_p), ("res_ld", C.c_int64), ("res_bstride", C.c_int64),
        ("scale", C.c_float), ("force_block_n", C.c_int), ("post_div", C.c_float), ("n_store", C.c_int), ("out2", C.c_void_p), ("force_mt", C.c_int), ("stats", C.c_void_p), ("force_halo", C.c_int), ("force_swap", C.c_int),
        ("gn_ab", C.c_void_p), ("gn_silu", C.c_int), ("poly", C.c_int),
    ]


class sdm_attn_args(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("heads", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int),
        ("q", C.c_void_p), ("ldq", C.c_int64), ("k", C.c_void_p), ("ldk", C.c_int64),
        ("vt", C.c_void_p), ("ldvt", C.c_int64), ("bias", C.c_void_p), ("bias_bstride", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("scale", C.c_float), ("ntiles", C.c_void_p),
    ]


class sdm_groupnorm_args(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("HW", C.c_int), ("nsrc", C.c_int),
        ("src0", C.c_void_p), ("c0", C.c_int), ("ld0", C.c_int64),
        ("src1", C.c_void_p), ("c1", C.c_int), ("ld1", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("silu", C.c_int),
        ("out", C.c_void_p), ("scratch", C.c_void_p), ("scratch_floats", C.c_size_t),
        ("pre0", C.c_void_p), ("pre1", C.c_void_p), ("pre_slots", C.c_int),
    ]


class sdm_direct_conv_args(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int), ("ksize", C.c_int),
        ("x", C.c_void_p), ("x_ld", C.c_int64), ("w", C.c_void_p), ("bias", C.c_void_p),
        ("out", C.c_void_p), ("out_ld", C.c_int64), ("out_coff", C.c_int), ("out_scale", C.c_float), ("cout_limit", C.c_int),
    ]


EXPORTS = [
    "sdm_version", "sdm_last_error", "sdm_create", "sdm_destroy", "sdm_load_weights", "sdm_load_report",
    "sdm_workspace_bytes", "sdm_forward", "sdm_workspace_bytes_prompt", "sdm_forward_prompt", "sdm_forward_host", "sdm_node_workspace_bytes", "sdm_apply_matte_host", "sdm_forward_profiled", "sdm_profile_count", "sdm_profile_entry",
    "sdm_last_forward_stats", "sdm_debug_tensor", "sdm_debug_tensor_count", "sdm_debug_tensor_name", "sdm_set_option", "sdm_graph_stats", "sdm_node_call_timing",
    "sdm_preprocess", "sdm_postprocess",
    "sdm_k_conv_gemm", "sdm_k_conv_tiles_per_image", "sdm_k_conv_variant", "sdm_k_conv_can_fuse_gn", "sdm_k_conv_can_poly", "sdm_k_groupnorm_ab_offset", "sdm_k_attention", "sdm_k_groupnorm_scratch_floats", "sdm_k_groupnorm", "sdm_k_layernorm",
    "sdm_k_softmax_rows", "sdm_k_direct_conv", "sdm_k_key_bias", "sdm_k_key_compact", "sdm_k_gather_rows", "sdm_k_probe_halo",
    "sdm_safetensors_open", "sdm_safetensors_count", "sdm_safetensors_entry", "sdm_safetensors_close",
]


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """dlopen the CUDA library (building it first if a toolkit is present). Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    # always go through build_ext.build(): it is gated by a digest of csrc/ + the public header, so it only compiles when the
    # sources changed — a stale .so from an older tree must never be loaded silently (struct layouts may have moved)
    import importlib.util

    spec = importlib.util.spec_from_file_location("_sdm_build_ext", os.path.join(_HERE, "build_ext.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    lib = C.CDLL(_LIB_PATH)
    lib.sdm_version.restype = C.c_int
    if lib.sdm_version() != ABI_VERSION:
        raise RuntimeError(f"libsdmatte_b200.so reports ABI {lib.sdm_version()}, this binding expects {ABI_VERSION}: rebuild "
                           f"(python comfyui-sdmatte_b200/build_ext.py --force)")
    lib.sdm_version.restype = C.c_int
    lib.sdm_last_error.restype = C.c_char_p
    lib.sdm_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.sdm_destroy.argtypes = [C.c_void_p]
    lib.sdm_destroy.restype = None
    lib.sdm_load_weights.argtypes = [C.c_void_p, C.POINTER(sdm_tensor_desc), C.c_int]
    lib.sdm_load_report.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.sdm_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.sdm_workspace_bytes.restype = C.c_size_t
    lib.sdm_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sdm_forward_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sdm_workspace_bytes_prompt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.sdm_workspace_bytes_prompt.restype = C.c_size_t
    lib.sdm_forward_prompt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sdm_node_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.sdm_node_workspace_bytes.restype = C.c_size_t
    lib.sdm_apply_matte_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                         C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sdm_preprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sdm_postprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sdm_forward_profiled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_void_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sdm_profile_count.argtypes = [C.c_void_p]
    lib.sdm_profile_entry.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]
    lib.sdm_last_forward_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    lib.sdm_debug_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    lib.sdm_debug_tensor_count.argtypes = [C.c_void_p]
    lib.sdm_debug_tensor_name.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    lib.sdm_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.sdm_graph_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.sdm_node_call_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.sdm_k_conv_gemm.argtypes = [C.POINTER(sdm_conv_gemm_args), C.c_void_p]
    lib.sdm_k_conv_tiles_per_image.argtypes = [C.c_int, C.c_int]
    lib.sdm_k_attention.argtypes = [C.POINTER(sdm_attn_args), C.c_void_p]
    lib.sdm_safetensors_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.sdm_safetensors_count.argtypes = [C.c_void_p]
    lib.sdm_safetensors_entry.argtypes = [C.c_void_p, C.c_int, C.POINTER(sdm_tensor_desc)]
    lib.sdm_safetensors_close.argtypes = [C.c_void_p]
    lib.sdm_safetensors_close.restype = None
    lib.sdm_k_key_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.sdm_k_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.sdm_k_groupnorm_scratch_floats.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.sdm_k_groupnorm_scratch_floats.restype = C.c_size_t
    lib.sdm_k_groupnorm.argtypes = [C.POINTER(sdm_groupnorm_args), C.c_void_p]
    lib.sdm_k_layernorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p]
    lib.sdm_k_softmax_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.sdm_k_direct_conv.argtypes = [C.POINTER(sdm_direct_conv_args), C.c_void_p]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise RuntimeError("sdmatte_b200: " + load_library().sdm_last_error().decode("utf-8", "replace"))


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class SafeTensorsReader:
    """Native .safetensors reader (csrc/safetensors.cu): header parse + read-only mmap, no torch tensors are created.
    `descs()` yields the sdm_tensor_desc array the engine's loader takes, filtered exactly like Engine.load_state_dict
    (unet.* / vae.* keys, floating-point dtypes, rank <= 4); the descriptors point into the mapping, so the reader must stay
    open until sdm_load_weights has returned."""

    def __init__(self, path: str):
        self.lib = load_library()
        self.h = C.c_void_p()
        _check(self.lib.sdm_safetensors_open(os.fsencode(path), C.byref(self.h)))

    def __len__(self):
        return self.lib.sdm_safetensors_count(self.h)

    def entry(self, i: int) -> "sdm_tensor_desc":
        d = sdm_tensor_desc()
        _check(self.lib.sdm_safetensors_entry(self.h, i, C.byref(d)))
        return d

    def descs(self):
        n_all = len(self)
        arr = (sdm_tensor_desc * max(1, n_all))()
        n = 0
        for i in range(n_all):
            d = self.entry(i)
            name = d.name.decode()
            if not (name.startswith("unet.") or name.startswith("vae.")) or d.dtype < 0 or d.ndim > 4:
                continue
            arr[n] = d
            n += 1
        return arr, n

    def close(self):
        if self.h:
            self.lib.sdm_safetensors_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Engine:
    """One handle per device. `forward` takes/returns CUDA tensors; `forward_host` takes/returns host tensors."""

    def __init__(self, device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("sdmatte_b200 needs a CUDA (sm_100a / B200) device; there is no CPU fallback")
        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib.sdm_create(C.byref(h), self.device.index or 0))
        self.h = h
        self._ws: Optional[torch.Tensor] = None
        self.load_report = (0, 0)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights (replaces load_state_dict + .to(device), sdmatte_nodes.py:298-323)
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        keep = []
        descs = (sdm_tensor_desc * len(sd))()
        n = 0
        for k, v in sd.items():
            if not (k.startswith("unet.") or k.startswith("vae.")):
                continue  # text_encoder.* is dead on this path
            if v.dtype not in _DTYPES:
                continue
            t = v.detach().to("cpu").contiguous()
            if t.dim() > 4:
                continue
            keep.append(t)
            d = descs[n]
            d.name = k.encode()
            d.dtype = _DTYPES[t.dtype]
            d.ndim = t.dim()
            for i, s in enumerate(t.shape):
                d.shape[i] = s
            d.data = t.data_ptr()
            n += 1
        _check(self.lib.sdm_load_weights(self.h, descs, n))
        used, unexpected = C.c_int(), C.c_int()
        _check(self.lib.sdm_load_report(self.h, C.byref(used), C.byref(unexpected)))
        self.load_report = (used.value, unexpected.value)
        self._ws = None
        return self.load_report

    def load_safetensors(self, path: str):
        """Load a checkpoint file through the native reader: the engine repacks straight from the file mapping
        (the reference builds torch CPU tensors with safe_open first, sdmatte_nodes.py:298-304)."""
        with SafeTensorsReader(path) as rd:
            descs, n = rd.descs()
            _check(self.lib.sdm_load_weights(self.h, descs, n))
        used, unexpected = C.c_int(), C.c_int()
        _check(self.lib.sdm_load_report(self.h, C.byref(used), C.byref(unexpected)))
        self.load_report = (used.value, unexpected.value)
        self._ws = None
        return self.load_report

    # ---- workspace (caller-owned arena: torch allocates, the engine never cudaMallocs per call)
    def workspace(self, B: int, R: int, host_staging: bool = False) -> torch.Tensor:
        need = self.lib.sdm_workspace_bytes(self.h, B, R)
        if need == 0:
            _check(1)
        if host_staging:
            need += B * R * R * (12 + 4 + 2) + 8192
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, image: torch.Tensor, trimap: torch.Tensor, is_transparent=False, want_premean: bool = False,
                out: Optional[torch.Tensor] = None):
        """image [B,R,R,3] fp32 cuda, trimap [B,R,R] fp32 cuda -> alpha [B,R,R] fp16 cuda (and pre-clip mean).
        `out`: optional preallocated alpha; calling again with the same input/output tensors replays the plan's CUDA graph."""
        B, R = image.shape[0], image.shape[1]
        assert image.shape == (B, R, R, 3) and trimap.shape == (B, R, R), "inputs must already be R x R"
        assert image.dtype == torch.float32 and trimap.dtype == torch.float32
        assert image.is_cuda and trimap.is_cuda
        image = image.contiguous()
        trimap = trimap.contiguous()
        ws = self.workspace(B, R)
        if out is not None:
            assert out.shape == (B, R, R) and out.dtype == torch.float16 and out.is_cuda and out.is_contiguous()
        alpha = out if out is not None else torch.empty((B, R, R), dtype=torch.float16, device=self.device)
        pre = torch.empty((B, R, R), dtype=torch.float16, device=self.device) if want_premean else None
        flags = is_transparent if isinstance(is_transparent, (list, tuple)) else [is_transparent] * B
        it = (C.c_int32 * B)(*[1 if f else 0 for f in flags])
        with torch.cuda.device(self.device):
            _check(self.lib.sdm_forward(self.h, image.data_ptr(), trimap.data_ptr(), B, R, it, alpha.data_ptr(),
                                        pre.data_ptr() if pre is not None else None, ws.data_ptr(), ws.numel(),
                                        _stream_ptr(self.device)))
        return (alpha, pre) if want_premean else alpha

    PROMPT_KINDS = {"trimap": 0, "mask": 0, "bbox_mask": 0, "point_mask": 1}

    def forward_prompt(self, image: torch.Tensor, aux: torch.Tensor, prompt: str, coords: torch.Tensor, is_transparent=False,
                       want_premean: bool = False):
        """The model's other visual prompts (include/sdmatte_b200.h: sdm_forward_prompt): `aux` [B,R,R] fp32 cuda is a mask, bbox
        mask or point mask, `coords` the matching [B,4] box / [B,N] point coordinates (host or device tensor, fp32)."""
        B, R = image.shape[0], image.shape[1]
        assert image.shape == (B, R, R, 3) and aux.shape == (B, R, R) and image.is_cuda and aux.is_cuda
        kind = self.PROMPT_KINDS[prompt]
        coords = coords.detach().to("cpu", torch.float32).contiguous()
        assert coords.dim() == 2 and coords.shape[0] == B
        n = coords.shape[1]
        image, aux = image.contiguous(), aux.contiguous()
        need = self.lib.sdm_workspace_bytes_prompt(self.h, B, R, kind, n)
        if need == 0:
            _check(1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        alpha = torch.empty((B, R, R), dtype=torch.float16, device=self.device)
        pre = torch.empty((B, R, R), dtype=torch.float16, device=self.device) if want_premean else None
        flags = is_transparent if isinstance(is_transparent, (list, tuple)) else [is_transparent] * B
        it = (C.c_int32 * B)(*[1 if f else 0 for f in flags])
        with torch.cuda.device(self.device):
            _check(self.lib.sdm_forward_prompt(self.h, image.data_ptr(), aux.data_ptr(), B, R, it, kind, coords.data_ptr(), n, alpha.data_ptr(),
                                               pre.data_ptr() if pre is not None else None, self._ws.data_ptr(), self._ws.numel(),
                                               _stream_ptr(self.device)))
        return (alpha, pre) if want_premean else alpha

    def forward_host(self, image: torch.Tensor, trimap: torch.Tensor, is_transparent=False, out: Optional[torch.Tensor] = None):
        """Host tensors in, host fp16 alpha out; H2D + D2H happen inside the library call (and it synchronises)."""
        B, R = image.shape[0], image.shape[1]
        assert image.shape == (B, R, R, 3) and trimap.shape == (B, R, R)
        assert not image.is_cuda and not trimap.is_cuda
        image = image.contiguous().float()
        trimap = trimap.contiguous().float()
        ws = self.workspace(B, R, host_staging=True)
        if out is None:
            out = torch.empty((B, R, R), dtype=torch.float16)
        flags = is_transparent if isinstance(is_transparent, (list, tuple)) else [is_transparent] * B
        it = (C.c_int32 * B)(*[1 if f else 0 for f in flags])
        with torch.cuda.device(self.device):
            _check(self.lib.sdm_forward_host(self.h, image.data_ptr(), trimap.data_ptr(), B, R, it, out.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        return out

    def apply_host(self, image: torch.Tensor, trimap: torch.Tensor, R: int, is_transparent=False, output_mode: str = "alpha_only",
                   mask_refine: bool = True, trimap_constraint: float = 0.8, alpha_out: Optional[torch.Tensor] = None,
                   matted_out: Optional[torch.Tensor] = None):
        """The node call (include/sdmatte_b200.h: sdm_apply_matte_host): HOST image [B,H,W,3] / trimap [B,H,W] fp32 of any size
        (pageable is fine) -> HOST alpha [B,H,W] fp16 and matted [B,H,W,3|4] fp32 (None for "alpha_only").  One library call:
        staging, H2D, resize, forward, post-processing, D2H, sync."""
        B, H, W, _ = image.shape
        assert image.shape == (B, H, W, 3) and trimap.shape == (B, H, W)
        assert not image.is_cuda and not trimap.is_cuda
        image = image.contiguous().float()
        trimap = trimap.contiguous().float()
        mode = OUTPUT_MODES.get(output_mode, 3)
        need = self.lib.sdm_node_workspace_bytes(self.h, B, H, W, int(R), mode)
        if need == 0:
            _check(1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if alpha_out is None:
            alpha_out = torch.empty((B, H, W), dtype=torch.float16)
        mch = 4 if mode == 1 else 3
        if mode != 0 and matted_out is None:
            matted_out = torch.empty((B, H, W, mch), dtype=torch.float32)
        flags = is_transparent if isinstance(is_transparent, (list, tuple)) else [is_transparent] * B
        it = (C.c_int32 * B)(*[1 if f else 0 for f in flags])
        with torch.cuda.device(self.device):
            _check(self.lib.sdm_apply_matte_host(self.h, image.data_ptr(), trimap.data_ptr(), B, H, W, int(R), it, int(bool(mask_refine)),
                                                 float(trimap_constraint), mode, alpha_out.data_ptr(),
                                                 matted_out.data_ptr() if mode != 0 else None, self._ws.data_ptr(), self._ws.numel(),
                                                 _stream_ptr(self.device)))
        return alpha_out, (matted_out if mode != 0 else None)

    def forward_profiled(self, image: torch.Tensor, trimap: torch.Tensor, is_transparent=False):
        """One forward with CUDA events around every op; returns [(kind, ms, flops, bytes), ...] (measurement aid)."""
        B, R = image.shape[0], image.shape[1]
        ws = self.workspace(B, R)
        alpha = torch.empty((B, R, R), dtype=torch.float16, device=self.device)
        it = (C.c_int32 * B)(*[1 if is_transparent else 0] * B)
        with torch.cuda.device(self.device):
            _check(self.lib.sdm_forward_profiled(self.h, image.data_ptr(), trimap.data_ptr(), B, R, it, alpha.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        out = []
        buf = C.create_string_buffer(64)
        ms, fl, by = C.c_float(), C.c_double(), C.c_double()
        for i in range(self.lib.sdm_profile_count(self.h)):
            _check(self.lib.sdm_profile_entry(self.h, i, buf, 64, C.byref(ms), C.byref(fl), C.byref(by)))
            out.append((buf.value.decode(), ms.value, fl.value, by.value))
        return out

    def stats(self):
        n, f = C.c_int(), C.c_double()
        _check(self.lib.sdm_last_forward_stats(self.h, C.byref(n), C.byref(f)))
        return {"launches": n.value, "tensor_flops": f.value}

    def set_option(self, name: str, value: int):
        """Diagnostics switches of the handle (include/sdmatte_b200.h: sdm_set_option); drops the cached plan and workspace."""
        _check(self.lib.sdm_set_option(self.h, name.encode(), int(value)))
        self._ws = None

    def graph_stats(self):
        c, l = C.c_int(), C.c_int()
        _check(self.lib.sdm_graph_stats(self.h, C.byref(c), C.byref(l)))
        return {"captures": c.value, "launches": l.value}

    def node_call_timing(self):
        """Host wall-clock split of the last apply_host call (ms): staging + H2D enqueue, kernel enqueue, GPU wait, copy-out."""
        ms = (C.c_double * 4)()
        _check(self.lib.sdm_node_call_timing(self.h, ms))
        return {"stage_in_ms": ms[0], "enqueue_ms": ms[1], "gpu_wait_ms": ms[2], "copy_out_ms": ms[3]}

    def tap_names(self):
        """Names of the block taps of the last forward, in graph order (all of them only with set_option("keep_taps", 1))."""
        buf = C.create_string_buffer(64)
        out = []
        for i in range(self.lib.sdm_debug_tensor_count(self.h)):
            _check(self.lib.sdm_debug_tensor_name(self.h, i, buf, 64))
            out.append(buf.value.decode())
        return out

    def debug_tensor(self, name: str) -> torch.Tensor:
        shape = (C.c_int64 * 4)()
        dt = C.c_int()
        _check(self.lib.sdm_debug_tensor(self.h, name.encode(), None, 0, shape, C.byref(dt)))
        out = torch.empty((shape[0], shape[1], shape[2], shape[3]), dtype=torch.float16, device=self.device)
        _check(self.lib.sdm_debug_tensor(self.h, name.encode(), out.data_ptr(), out.numel() * 2, shape, C.byref(dt)))
        return out


# ---------------------------------------------------------------------------------------------------------------
# node-side pre/post-processing on the device (reference sdmatte_nodes.py:204-214,339-397)
# ---------------------------------------------------------------------------------------------------------------
OUTPUT_MODES = {"alpha_only": 0, "matted_rgba": 1, "matted_rgb": 2}


def preprocess(image: torch.Tensor, trimap: torch.Tensor, R: int):
    """image [B,H,W,3] fp32 cuda, trimap [B,H,W] fp32 cuda -> antialias-bilinear resized ([B,R,R,3], [B,R,R]) fp32 cuda
    (torchvision Resize(antialias=True) semantics).  Identity when the inputs already are R x R."""
    B, H, W, _ = image.shape
    assert image.is_cuda and trimap.is_cuda and image.dtype == torch.float32 and trimap.dtype == torch.float32
    assert trimap.shape == (B, H, W)
    image, trimap = image.contiguous(), trimap.contiguous()
    if (H, W) == (R, R):
        return image, trimap
    img_r = torch.empty((B, R, R, 3), dtype=torch.float32, device=image.device)
    tri_r = torch.empty((B, R, R), dtype=torch.float32, device=image.device)
    with torch.cuda.device(image.device):
        _check(load_library().sdm_preprocess(image.data_ptr(), trimap.data_ptr(), B, H, W, R, img_r.data_ptr(), tri_r.data_ptr(),
                                             _stream_ptr(image.device)))
    return img_r, tri_r


def postprocess(alpha: torch.Tensor, image: torch.Tensor, trimap: torch.Tensor, output_mode: str, mask_refine: bool,
                trimap_constraint: float):
    """alpha [B,R,R] fp16 cuda (engine output), image [B,H,W,3] / trimap [B,H,W] fp32 cuda (the caller's originals) ->
    (alpha_out [B,H,W] fp16 cuda, matted [B,H,W,3|4] fp32 cuda or None for "alpha_only")."""
    B, R = alpha.shape[0], alpha.shape[1]
    _, H, W, _ = image.shape
    assert alpha.is_cuda and alpha.dtype == torch.float16 and alpha.shape == (B, R, R)
    mode = OUTPUT_MODES.get(output_mode, 3)
    alpha, image, trimap = alpha.contiguous(), image.contiguous(), trimap.contiguous()
    out = torch.empty((B, H, W), dtype=torch.float16, device=alpha.device)
    matted = None
    if mode != 0:
        matted = torch.empty((B, H, W, 4 if mode == 1 else 3), dtype=torch.float32, device=alpha.device)
    with torch.cuda.device(alpha.device):
        _check(load_library().sdm_postprocess(alpha.data_ptr(), B, R, H, W, image.data_ptr(), trimap.data_ptr(), int(bool(mask_refine)),
                                              float(trimap_constraint), mode, out.data_ptr(),
                                              matted.data_ptr() if matted is not None else None, _stream_ptr(alpha.device)))
    return out, matted


# ---------------------------------------------------------------------------------------------------------------
# kernel-level entry points (used by tests/bench only)
# ---------------------------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


def k_conv_gemm(srcs, w, N, out, *, B, Hin, Win, ksize=1, stride=1, pad=0, mode=0, ups2=0, bias=None, bias_sel=None,
                res=None, scale=1.0, w_bstride=0, out_ld=None, out_bstride=None, force_block_n=0, post_div=1.0, n_store=0, out2=None, force_mt=0, stats=None, force_halo=0, force_swap=0, gn_ab=None, gn_silu=0, poly=0):
    lib = load_library()
    a = sdm_conv_gemm_args()
    a.poly = int(poly)
    a.B, a.Hin, a.Win, a.nsrc = B, Hin, Win, len(srcs)
    a.src0, a.c0, a.ld0 = srcs[0][0].data_ptr(), srcs[0][1], srcs[0][2]
    if len(srcs) > 1:
        a.src1, a.c1, a.ld1 = srcs[1][0].data_ptr(), srcs[1][1], srcs[1][2]
    a.ksize, a.stride, a.pad = ksize, stride, pad
    a.w, a.N, a.w_bstride = w.data_ptr(), N, w_bstride
    a.mode, a.ups2 = mode, ups2
    a.out, a.out_ld, a.out_bstride = out.data_ptr(), out_ld, out_bstride
    a.bias, a.bias_sel = _p(bias), _p(bias_sel)
    if res is not None:
        a.res, a.res_ld, a.res_bstride = res[0].data_ptr(), res[1], res[2]
    a.scale, a.force_block_n = scale, force_block_n
    a.post_div, a.n_store, a.out2, a.force_mt, a.stats = post_div, n_store, _p(out2), force_mt, _p(stats)
    a.force_halo, a.force_swap = force_halo, force_swap
    a.gn_ab, a.gn_silu = _p(gn_ab), int(gn_silu)
    _check(lib.sdm_k_conv_gemm(C.byref(a), _stream_ptr(out.device)))


def k_attention(q, k, vt, out, *, B, heads, Lq, Lk, ldq, ldk, ldvt, ldo, bias=None, bias_bstride=0, scale=0.125, ntiles=None):
    lib = load_library()
    a = sdm_attn_args()
    a.B, a.heads, a.Lq, a.Lk = B, heads, Lq, Lk
    a.q, a.ldq, a.k, a.ldk, a.vt, a.ldvt = q.data_ptr(), ldq, k.data_ptr(), ldk, vt.data_ptr(), ldvt
    a.bias, a.bias_bstride = _p(bias), bias_bstride
    a.out, a.ldo, a.scale = out.data_ptr(), ldo, scale
    a.ntiles = _p(ntiles)
    _check(lib.sdm_k_attention(C.byref(a), _stream_ptr(out.device)))


def k_key_bias(trimap, R):
    """trimap [B,R,R] fp32 cuda -> the four per-level key-bias tensors [B][lpad_l] (log2 domain, -inf padded to 128 keys)."""
    lib = load_library()
    lib.sdm_k_key_bias.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    B = trimap.shape[0]
    S = R // 8
    lpad = [(((S >> l) ** 2 + 127) // 128) * 128 for l in range(4)]
    outs = [torch.empty((B, lp), dtype=torch.float32, device=trimap.device) for lp in lpad]
    arr = (C.c_int32 * 4)(*lpad)
    _check(lib.sdm_k_key_bias(trimap.contiguous().data_ptr(), B, R, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(),
                              arr, _stream_ptr(trimap.device)))
    return outs


def k_key_compact(bias, cbias, idx, ntiles, *, B, L, lpad):
    """bias/cbias float32 [B][lpad], idx int32 [B][lpad], ntiles int32 [B] (see include/sdmatte_b200.h)."""
    _check(load_library().sdm_k_key_compact(bias.data_ptr(), cbias.data_ptr(), idx.data_ptr(), ntiles.data_ptr(), B, L, lpad,
                                           _stream_ptr(bias.device)))


def k_probe_halo(x, eye, out, dy, dx, mode):
    lib = load_library()
    lib.sdm_k_probe_halo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    _check(lib.sdm_k_probe_halo(x.data_ptr(), eye.data_ptr(), out.data_ptr(), dy, dx, mode, _stream_ptr(x.device)))


def k_gather_rows(src, dst, idx, ntiles, *, B, L, C_, idx_bstride):
    _check(load_library().sdm_k_gather_rows(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), ntiles.data_ptr(), B, L, C_, idx_bstride,
                                           _stream_ptr(src.device)))


def conv_variant(ksize, stride, N, H, W, *, mode=0, ups2=0, has_res=0):
    """Kernel variant for a conv of this per-sample geometry: 0 tap-per-box, 1 / 2 halo (256 / 160 wide), 3 swapped operands."""
    lib = load_library()
    lib.sdm_k_conv_variant.argtypes = [C.c_int] * 8
    return lib.sdm_k_conv_variant(ksize, stride, mode, ups2, N, has_res, H, W)


def conv_can_fuse_gn(ksize, stride, N, H, W, *, mode=0, ups2=0, has_res=0):
    lib = load_library()
    lib.sdm_k_conv_can_fuse_gn.argtypes = [C.c_int] * 8
    return bool(lib.sdm_k_conv_can_fuse_gn(ksize, stride, mode, ups2, N, has_res, H, W))


def conv_can_poly(N, H, W):
    """Can "nearest x2 upsample -> 3x3 conv to N channels" of an H x W input run as four polyphase launches (k_conv_gemm(poly=1..4))?"""
    lib = load_library()
    lib.sdm_k_conv_can_poly.argtypes = [C.c_int] * 3
    return bool(lib.sdm_k_conv_can_poly(N, H, W))


def conv_tiles_per_image(H, W):
    return load_library().sdm_k_conv_tiles_per_image(H, W)


def k_groupnorm(srcs, gamma, beta, out, *, B, HW, eps, silu, pre=None, pre_slots=0):
    lib = load_library()
    a = sdm_groupnorm_args()
    a.B, a.HW, a.nsrc = B, HW, len(srcs)
    a.src0, a.c0, a.ld0 = srcs[0][0].data_ptr(), srcs[0][1], srcs[0][2]
    ctot = srcs[0][1]
    if len(srcs) > 1:
        a.src1, a.c1, a.ld1 = srcs[1][0].data_ptr(), srcs[1][1], srcs[1][2]
        ctot += srcs[1][1]
    n = lib.sdm_k_groupnorm_scratch_floats(B, HW, ctot)
    dev = srcs[0][0].device
    scratch = torch.empty(n, dtype=torch.float32, device=dev)
    a.gamma, a.beta, a.eps, a.silu = gamma.data_ptr(), beta.data_ptr(), eps, int(silu)
    a.out, a.scratch, a.scratch_floats = _p(out), scratch.data_ptr(), n
    if pre is not None:
        a.pre0, a.pre1, a.pre_slots = _p(pre[0]), _p(pre[1]) if len(pre) > 1 else None, pre_slots
    _check(lib.sdm_k_groupnorm(C.byref(a), _stream_ptr(dev)))
    return scratch


def groupnorm_ab(scratch, B, HW, Ctot):
    """The [B][Ctot][2] (scale, shift) table k_groupnorm left in its scratch buffer (view, no copy)."""
    lib = load_library()
    lib.sdm_k_groupnorm_ab_offset.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.sdm_k_groupnorm_ab_offset.restype = C.c_size_t
    off = lib.sdm_k_groupnorm_ab_offset(B, HW, Ctot)
    return scratch[off: off + B * Ctot * 2]


def k_layernorm(x, y, gamma, beta, rows, Cc, eps=1e-5):
    _check(load_library().sdm_k_layernorm(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, Cc, eps,
                                          _stream_ptr(x.device)))


def k_softmax_rows(s, p, rows, L):
    _check(load_library().sdm_k_softmax_rows(s.data_ptr(), p.data_ptr(), rows, L, _stream_ptr(s.device)))


def k_direct_conv(x, w, bias, out, *, B, H, W, Cin, Cout, ksize, x_ld, out_ld, out_coff=0, out_scale=1.0, cout_limit=0):
    a = sdm_direct_conv_args()
    a.B, a.H, a.W, a.Cin, a.Cout, a.ksize = B, H, W, Cin, Cout, ksize
    a.x, a.x_ld, a.w, a.bias = x.data_ptr(), x_ld, w.data_ptr(), _p(bias)
    a.out, a.out_ld, a.out_coff, a.out_scale, a.cout_limit = out.data_ptr(), out_ld, out_coff, out_scale, cout_limit
    _check(load_library().sdm_k_direct_conv(C.byref(a), _stream_ptr(out.device)))
