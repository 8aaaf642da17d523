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


@pytest.mark.parametrize("case", ["deep_nesting", "metadata_not_flat", "uint_overflow", "numel_overflow", "duplicate_name", "overlap", "gap",
                                  "trailing_bytes", "huge_shape"])
def test_reader_is_hardened_against_malformed_headers(pkg, tmp_path, case):
    """Checkpoint files are untrusted input: bounded recursion, overflow-checked integers and products, and the
    safetensors package's own layout rules (no duplicate names, tensors tile the byte buffer exactly)."""
    E = pkg.engine
    path = str(tmp_path / f"{case}.safetensors")
    w = {"dtype": "F32", "shape": [2], "data_offsets": [0, 8]}
    payload = b"\0" * 16
    raw = None
    if case == "deep_nesting":  # 4 MB of '[' inside an unknown per-tensor key: must fail cleanly, not overflow the stack
        raw = b'{"unet.w":{"dtype":"F32","shape":[2],"data_offsets":[0,8],"x":' + b"[" * (4 << 20) + b"}}"
        payload = b"\0" * 8
    elif case == "metadata_not_flat":
        raw = b'{"__metadata__":{"a":{"b":"c"}},"unet.w":{"dtype":"F32","shape":[2],"data_offsets":[0,8]}}'
        payload = b"\0" * 8
    elif case == "uint_overflow":
        raw = b'{"unet.w":{"dtype":"F32","shape":[2],"data_offsets":[0,99999999999999999999999]}}'
    elif case == "numel_overflow":  # 2^32 * 2^32 wraps to 0 in 64 bits
        raw = b'{"unet.w":{"dtype":"F32","shape":[4294967296,4294967296],"data_offsets":[0,0]}}'
        payload = b""
    elif case == "huge_shape":
        raw = b'{"unet.w":{"dtype":"F32","shape":[9223372036854775807,3],"data_offsets":[0,8]}}'
        payload = b"\0" * 8
    elif case == "duplicate_name":
        raw = b'{"unet.w":{"dtype":"F32","shape":[2],"data_offsets":[0,8]},"unet.w":{"dtype":"F32","shape":[2],"data_offsets":[8,16]}}'
    elif case == "overlap":
        hdr = {"unet.a": w, "unet.b": {"dtype": "F32", "shape": [2], "data_offsets": [4, 12]}}
        payload = b"\0" * 12
    elif case == "gap":
        hdr = {"unet.a": w, "unet.b": {"dtype": "F32", "shape": [1], "data_offsets": [12, 16]}}
    else:  # trailing_bytes
        hdr = {"unet.a": w}
    if raw is not None:
        with open(path, "wb") as f:
            f.write(struct.pack("<Q", len(raw)) + raw + payload)
    else:
        _write_raw(path, hdr, payload)
    with pytest.raises(RuntimeError):
        E.SafeTensorsReader(path)
    if case in ("duplicate_name", "overlap", "gap", "trailing_bytes", "metadata_not_flat"):  # ...and the package agrees
        with pytest.raises(Exception):
            safetensors_torch.load_file(path)


def test_reader_accepts_flat_metadata_and_empty_tensor(pkg, tmp_path):
    E = pkg.engine
    path = str(tmp_path / "ok.safetensors")
    _write_raw(path, {"__metadata__": {"format": "pt"}, "unet.e": {"dtype": "F32", "shape": [0, 4], "data_offsets": [0, 0]},
                      "unet.w": {"dtype": "F16", "shape": [2, 2], "data_offsets": [0, 8]}}, b"\1" * 8)
    with E.SafeTensorsReader(path) as rd:
        assert len(rd) == 2
        for i in range(2):
            rd.entry(i)
    assert set(safetensors_torch.load_file(path)) == {"unet.e", "unet.w"}
