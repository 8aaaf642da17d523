"""CPU suite for the native .safetensors reader (csrc/safetensors.cu): every tensor the engine's loader would
see through the native reader is byte-identical to what the `safetensors` package returns (the reader the reference uses,
sdmatte_nodes.py:298-304), including unaligned data sections, metadata, non-float and high-rank tensors, and broken files."""
import ctypes as C
import json
import os
import struct

import pytest
import torch

safetensors_torch = pytest.importorskip("safetensors.torch")


def _bytes_of(desc, nbytes):
    return C.string_at(desc.data, nbytes)


def _write_raw(path, header: dict, payload: bytes, pad_to: int = 1):
    h = json.dumps(header, separators=(",", ":")).encode()
    while (8 + len(h)) % pad_to:
        h += b" "
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        f.write(payload)


def test_reader_matches_safetensors_package(pkg, tmp_path):
    E = pkg.engine
    g = torch.Generator().manual_seed(0)
    sd = {
        "unet.conv_in.weight": torch.randn(320, 8, 3, 3, generator=g),
        "unet.conv_in.bias": torch.randn(320, generator=g).half(),
        "vae.decoder.conv_out.weight": torch.randn(3, 128, 3, 3, generator=g).bfloat16(),
        "vae.quant_conv.weight": torch.randn(8, 8, 1, 1, generator=g),
        "text_encoder.embeddings.weight": torch.randn(7, 5, generator=g),      # dead on this path: filtered by prefix
        "unet.step": torch.tensor([3], dtype=torch.int64),                      # non-float: filtered by dtype
        "unet.rank5": torch.zeros(1, 2, 1, 2, 3),                               # rank > 4: filtered
    }
    path = str(tmp_path / "ckpt.safetensors")
    safetensors_torch.save_file(sd, path, metadata={"format": "pt", "note": "quotes \" and \\ backslashes, {braces} [brackets]"})
    with E.SafeTensorsReader(path) as rd:
        assert len(rd) == len(sd)
        seen = {}
        for i in range(len(rd)):
            d = rd.entry(i)
            name = d.name.decode()
            t = sd[name]
            assert d.ndim == t.dim()
            assert list(d.shape[:min(4, t.dim())]) == list(t.shape[:4])
            code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}.get(t.dtype, -1)
            assert d.dtype == code
            assert _bytes_of(d, t.numel() * t.element_size()) == t.contiguous().view(torch.uint8).numpy().tobytes(), name
            seen[name] = d
        assert set(seen) == set(sd)
        arr, n = rd.descs()
        kept = sorted(arr[i].name.decode() for i in range(n))
        assert kept == ["unet.conv_in.bias", "unet.conv_in.weight", "vae.decoder.conv_out.weight", "vae.quant_conv.weight"]


def test_reader_handles_unaligned_data_section(pkg, tmp_path):
    """Header lengths that leave the data section at an odd address: fp32 / fp16 tensors come back through an aligned copy."""
    E = pkg.engine
    a = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    b = torch.arange(4, dtype=torch.float16)
    payload = a.numpy().tobytes() + b.numpy().tobytes()
    header = {"vae.a": {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 24]},
              "vae.b": {"dtype": "F16", "shape": [4], "data_offsets": [24, 32]}}
    for pad in (1, 2, 3, 5, 8):
        path = str(tmp_path / f"u{pad}.safetensors")
        _write_raw(path, header, payload, pad_to=pad)
        with E.SafeTensorsReader(path) as rd:
            got = {rd.entry(i).name.decode(): rd.entry(i) for i in range(len(rd))}
            assert got["vae.a"].data % 4 == 0 and got["vae.b"].data % 2 == 0
            assert _bytes_of(got["vae.a"], 24) == a.numpy().tobytes()
            assert _bytes_of(got["vae.b"], 8) == b.numpy().tobytes()


@pytest.mark.parametrize("case", ["truncated", "offsets_outside", "size_mismatch", "not_json", "missing"])
def test_reader_rejects_broken_files(pkg, tmp_path, case):
    E = pkg.engine
    path = str(tmp_path / f"{case}.safetensors")
    good = {"unet.w": {"dtype": "F32", "shape": [2], "data_offsets": [0, 8]}}
    if case == "truncated":
        open(path, "wb").write(b"\x10\x00\x00")
    elif case == "offsets_outside":
        _write_raw(path, {"unet.w": {"dtype": "F32", "shape": [2], "data_offsets": [0, 64]}}, b"\0" * 8)
    elif case == "size_mismatch":
        _write_raw(path, {"unet.w": {"dtype": "F32", "shape": [3], "data_offsets": [0, 8]}}, b"\0" * 8)
    elif case == "not_json":
        with open(path, "wb") as f:
            f.write(struct.pack("<Q", 5) + b"hello" + b"\0" * 8)
    else:
        path = str(tmp_path / "does_not_exist.safetensors")
    del good
    if case == "size_mismatch":  # the header parses; the inconsistency is reported when the entry is requested
        with E.SafeTensorsReader(path) as rd:
            with pytest.raises(RuntimeError):
                rd.entry(0)
    else:
        with pytest.raises(RuntimeError):
            E.SafeTensorsReader(path)
