"""GPU probe (not a test): shifted windows of a TMA-swizzled halo tile as tcgen05 A operands.  Prints, per (dy, dx) and
descriptor mode, whether D equals the expected window.  python tests/probe_halo.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

E = ge.load_package().engine
torch.manual_seed(0)
x = torch.randn(16, 8, 64, device="cuda").half()
eye = torch.eye(64, device="cuda").half()
pad = torch.zeros(18, 10, 64, device="cuda")
pad[1:17, 1:9] = x.float()
for mode in (0, 1):
    ok = 0
    for dy in range(3):
        for dx in range(3):
            out = torch.full((128, 64), float("nan"), device="cuda")
            E.k_probe_halo(x, eye, out, dy, dx, mode)
            torch.cuda.synchronize()
            want = pad[dy:dy + 16, dx:dx + 8].reshape(128, 64)
            err = (out - want).abs().max().item()
            rows_ok = int(((out - want).abs().amax(1) == 0).sum())
            print(f"mode {mode} dy {dy} dx {dx}: max err {err:.3e}, exact rows {rows_ok}/128")
            ok += err == 0
    print(f"mode {mode}: {ok}/9 windows exact")
