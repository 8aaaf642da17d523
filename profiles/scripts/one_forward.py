"""ONE forward of the headline configuration (bs=8, 1024^2) for profiler runs: weights + plan are built, one eager warm-up forward
sets the kernel attributes, then exactly one more forward runs (eager: every launch visible to ncu).  SDM_ONE_B / SDM_ONE_R override."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
from oracle import synth  # noqa: E402

B, R = int(os.environ.get("SDM_ONE_B", "8")), int(os.environ.get("SDM_ONE_R", "1024"))
pkg = ge.load_package()
eng = pkg.engine.Engine(0)
eng.load_state_dict(synth.make_checkpoint(seed=1234))
eng.set_option("cuda_graph", 0)
image, trimap = synth.make_inputs(B, R, seed=1000)
img, tri = image.cuda(), trimap.cuda()
out = torch.empty((B, R, R), dtype=torch.float16, device="cuda")
if os.environ.get("SDM_ONE_WARM", "1") == "1":
    eng.forward(img, tri, False, out=out)
    torch.cuda.synchronize()
torch.cuda.nvtx.range_push("one_forward")
eng.forward(img, tri, False, out=out)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("launches", eng.stats()["launches"])
