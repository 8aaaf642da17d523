"""Checkpoint bootstrap for the `Apply SDMatte` node (reference: sdmatte_nodes.py:9-199).

What the reference does before it can run: (1) `download_model` — look for the checkpoint in every registered "SDMatte"
model folder, else fetch it from Hugging Face into `models/SDMatte/` through a `.tmp` file and an atomic rename
(sdmatte_nodes.py:103-199); (2) `ensure_sd21_from_manojb` — fetch ten Stable-Diffusion-2.1 config JSON / tokenizer files so that
diffusers can BUILD the model classes (sdmatte_nodes.py:33-101).  Step (2) has no counterpart here on purpose: the engine's
architecture is compiled in (SD-2.1 CustomUNet + SD VAE, csrc/engine.cu), it reads no config file and no tokenizer (the CLIP text
encoder is dead compute on this path).  Step (1) is kept with the same search order, file names, URLs and failure modes, with
the transport injectable so that it is testable offline.
"""
from __future__ import annotations

import os
import tempfile
from typing import Callable, Dict, Iterable, Optional

MODEL_URLS: Dict[str, str] = {
    "SDMatte.safetensors": "https://huggingface.co/1038lab/SDMatte/resolve/main/SDMatte.safetensors",
    "SDMatte_plus.safetensors": "https://huggingface.co/1038lab/SDMatte/resolve/main/SDMatte_plus.safetensors",
}

# fetch(url, destination_path) -> expected size in bytes (0 when the server did not say)
Fetcher = Callable[[str, str], int]


def _nonempty_file(path: str) -> bool:
    try:
        return os.path.isfile(path) and os.path.getsize(path) > 0
    except OSError:
        return False


def _default_fetch(url: str, dst: str) -> int:
    """requests (streamed, 1 MiB chunks, optional tqdm bar) when importable, else urllib — the reference's two transports."""
    try:
        import requests
    except ImportError:
        import urllib.request

        urllib.request.urlretrieve(url, dst)
        return 0
    try:
        from tqdm import tqdm
    except Exception:
        tqdm = None
    with requests.get(url, stream=True, timeout=60) as resp:
        resp.raise_for_status()
        total = int(resp.headers.get("content-length", 0) or 0)
        bar = tqdm(desc=os.path.basename(dst), total=total, unit="iB", unit_scale=True, unit_divisor=1024) if tqdm and total else None
        with open(dst, "wb") as f:
            for chunk in resp.iter_content(chunk_size=1 << 20):
                if chunk:
                    f.write(chunk)
                    if bar:
                        bar.update(len(chunk))
        if bar:
            bar.close()
    return total


def locate(model_name: str, search_paths: Iterable[str], models_dir: str) -> Optional[str]:
    """First non-empty `<folder>/<model_name>` over the registered SDMatte folders, then `models_dir` (sdmatte_nodes.py:104-130)."""
    for folder in list(search_paths) + [models_dir]:
        p = os.path.join(folder, model_name)
        if _nonempty_file(p):
            return p
    return None


def download_model(model_name: str, models_dir: str, search_paths: Iterable[str] = (), model_urls: Optional[Dict[str, str]] = None,
                   fetch: Optional[Fetcher] = None, offline: Optional[bool] = None) -> str:
    """Path of the checkpoint, downloading it if it is nowhere on disk.

    Same contract as the reference's download_model: ValueError for a name without a URL, the file lands in `models_dir` via
    `<name>.tmp` + atomic rename, a short download (size != content-length) is an IOError and leaves no partial file, and a
    concurrent process that finished first wins.  `offline` (default: env SDMATTE_OFFLINE=1) turns the download into a
    FileNotFoundError that says where the file is expected — the right behaviour on an air-gapped GPU box.
    """
    urls = MODEL_URLS if model_urls is None else model_urls
    found = locate(model_name, search_paths, models_dir)
    if found:
        return found
    url = urls.get(model_name)
    if not url:
        raise ValueError(f"[SDMatte] Unknown model name: {model_name}")
    target = os.path.join(models_dir, model_name)
    if offline is None:
        offline = os.environ.get("SDMATTE_OFFLINE", "0") == "1"
    if offline:
        raise FileNotFoundError(f"[SDMatte] '{model_name}' not found in {list(search_paths) + [models_dir]} and downloads are disabled "
                                f"(SDMATTE_OFFLINE=1); place the file at {target}")
    os.makedirs(models_dir, exist_ok=True)
    fd, tmp = tempfile.mkstemp(prefix=model_name + ".", suffix=".tmp", dir=models_dir)  # per-process name: no clobbering
    os.close(fd)
    try:
        print(f"[SDMatte] Model '{model_name}' not found. Downloading to {target}...")
        expected = (fetch or _default_fetch)(url, tmp)
        got = os.path.getsize(tmp)
        if got == 0 or (expected and got != expected):
            raise IOError(f"[SDMatte] Incomplete download: {got} != {expected}")
        if _nonempty_file(target):  # another process finished first
            return target
        os.replace(tmp, target)
        tmp = None
        return target
    finally:
        if tmp and os.path.exists(tmp):
            try:
                os.remove(tmp)
            except OSError:
                pass
