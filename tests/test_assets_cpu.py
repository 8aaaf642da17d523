"""Checkpoint bootstrap: the search order / download contract of the reference's download_model
(sdmatte_nodes.py:103-199), exercised offline with an injected transport and with a file:// URL through the urllib path."""
import os

import pytest


@pytest.fixture()
def assets(pkg):
    import importlib

    return importlib.import_module(pkg.__name__ + ".assets")


def test_search_order_registered_folders_first(assets, tmp_path):
    a, b, models = tmp_path / "a", tmp_path / "b", tmp_path / "models"
    for d in (a, b, models):
        d.mkdir()
    (b / "SDMatte.safetensors").write_bytes(b"B")
    (models / "SDMatte.safetensors").write_bytes(b"M")
    (a / "SDMatte.safetensors").write_bytes(b"")  # empty files are skipped, like the reference's getsize() > 0 test
    calls = []
    got = assets.download_model("SDMatte.safetensors", str(models), [str(a), str(b)], fetch=lambda u, d: calls.append(u) or 0)
    assert got == str(b / "SDMatte.safetensors") and not calls
    (b / "SDMatte.safetensors").unlink()
    assert assets.download_model("SDMatte.safetensors", str(models), [str(a), str(b)]) == str(models / "SDMatte.safetensors")


def test_download_is_atomic_and_size_checked(assets, tmp_path):
    models = tmp_path / "models"
    payload = os.urandom(4096)

    def good(url, dst):
        assert url == assets.MODEL_URLS["SDMatte_plus.safetensors"]
        open(dst, "wb").write(payload)
        return len(payload)

    p = assets.download_model("SDMatte_plus.safetensors", str(models), [], fetch=good, offline=False)
    assert p == str(models / "SDMatte_plus.safetensors") and open(p, "rb").read() == payload
    assert os.listdir(models) == ["SDMatte_plus.safetensors"]  # no .tmp left behind

    def short(url, dst):
        open(dst, "wb").write(payload[:100])
        return len(payload)

    with pytest.raises(IOError):
        assets.download_model("SDMatte.safetensors", str(models), [], fetch=short, offline=False)
    assert os.listdir(models) == ["SDMatte_plus.safetensors"]

    def boom(url, dst):
        open(dst, "wb").write(b"partial")
        raise ConnectionError("network down")

    with pytest.raises(ConnectionError):
        assets.download_model("SDMatte.safetensors", str(models), [], fetch=boom, offline=False)
    assert os.listdir(models) == ["SDMatte_plus.safetensors"]


def test_concurrent_winner_and_unknown_name_and_offline(assets, tmp_path):
    models = tmp_path / "models"

    def racing(url, dst):  # another process finishes while we download
        open(dst, "wb").write(b"ours")
        open(os.path.join(str(models), "SDMatte.safetensors"), "wb").write(b"theirs")
        return 4

    p = assets.download_model("SDMatte.safetensors", str(models), [], fetch=racing, offline=False)
    assert open(p, "rb").read() == b"theirs" and os.listdir(models) == ["SDMatte.safetensors"]
    with pytest.raises(ValueError):
        assets.download_model("other.safetensors", str(models), [], fetch=racing, offline=False)
    with pytest.raises(FileNotFoundError) as e:
        assets.download_model("SDMatte_plus.safetensors", str(models), [], offline=True)
    assert "SDMatte_plus.safetensors" in str(e.value)


def test_urllib_transport_with_file_url(assets, tmp_path, monkeypatch):
    """The reference falls back to urllib when `requests` is missing; a file:// URL drives that transport without a network."""
    import builtins

    src = tmp_path / "remote.bin"
    src.write_bytes(b"x" * 1000)
    real_import = builtins.__import__

    def no_requests(name, *a, **k):
        if name == "requests":
            raise ImportError("requests hidden for this test")
        return real_import(name, *a, **k)

    monkeypatch.setattr(builtins, "__import__", no_requests)
    p = assets.download_model("m.safetensors", str(tmp_path / "models"), [], model_urls={"m.safetensors": src.as_uri()}, offline=False)
    assert open(p, "rb").read() == b"x" * 1000
