import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
from oracle import synth
pkg = ge.load_package(); nodes = pkg.sdmatte_nodes
nodes.register_state_dict("SDMatte.safetensors", synth.make_checkpoint(seed=1234))
dev = torch.device("cuda", 0); nodes.set_devices([dev])
eng = nodes.get_engine("SDMatte.safetensors", dev)
B, R = 8, 1024
image, trimap = synth.make_inputs(B, R, seed=1000)
node = nodes.SDMatteApply()
img_d, tri_d = image.cuda(), trimap.cuda(); out = torch.empty((B, R, R), dtype=torch.float16, device=dev)
def dev_step():
    eng.forward(img_d, tri_d, False, out=out)
for _ in range(3): dev_step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): dev_step()
torch.cuda.synchronize(); print("device forward ms", (time.perf_counter() - t0) / 5 * 1e3)
for thr in (8, 4, 16, 12):
    eng.set_option("copy_threads", thr)
    for mode in ("alpha_only", "matted_rgba"):
        for _ in range(2): node.apply_matte("SDMatte.safetensors", image, trimap, R, False, mode, True, 0.8)
        t0 = time.perf_counter()
        for _ in range(4): node.apply_matte("SDMatte.safetensors", image, trimap, R, False, mode, True, 0.8)
        dt = (time.perf_counter() - t0) / 4 * 1e3
        print(f"threads {thr} {mode}: node call {dt:.2f} ms  split {eng.node_call_timing()}")
# direct apply_host with preallocated outputs (no python-side allocation)
a = torch.empty((B, R, R), dtype=torch.float16)
for _ in range(2): eng.apply_host(image, trimap, R, False, "alpha_only", True, 0.8, alpha_out=a)
t0 = time.perf_counter()
for _ in range(4): eng.apply_host(image, trimap, R, False, "alpha_only", True, 0.8, alpha_out=a)
print("apply_host direct", (time.perf_counter() - t0) / 4 * 1e3, eng.node_call_timing())
