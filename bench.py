#!/usr/bin/env python
"""bench.py — mattes/sec of the SDMatte single-pass matte path at 1024^2, bs=8 per GPU (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]                  # B200 engine (this repo)
  python bench.py --impl reference [--gpus N] --steps K --warmup W     # the reference graph on the host CPU (oracle port)
  python bench.py --size 768 --batch 8 / --batch 1 / --trimap allfg / --graph 0      # other BASELINE configs, A/B switches

One "step" = one pass of the hot path over one batch of `--batch` synthetic `--size`^2 RGB+trimap inputs per GPU.
N > 1: launched by torchrun, one rank per GPU, batch-sharded (weak scaling: `--batch` mattes per GPU), one NCCL all-gather of
the fp16 alpha per step (north_star); timing = CUDA events, barrier + synchronize on both sides, max over ranks.

JSON keys beyond the base contract:
  roofline      : the dominant kernel family, the tcgen05 3x3 convolutions (half of the step): achieved = algorithmic FLOPs of
                  those launches / their CUDA-event time inside one profiled step; peak = MEASURED_PEAKS.json bf16 sustained;
                  traffic = ncu dram bytes per launch of the same family (profiles/r2_traffic.json, captured by
                  profiles/scripts/run_r2_traffic.sh) or null.
  gemm_roofline : same for all non-attention tcgen05 kernels (round 1's `roofline`).
  path_roofline : whole step: B * TFLOP(R) / step time / sustained peak.
  kernel_breakdown: per kernel family ms / share / achieved TFLOP/s or GB/s (from the same profiled step).
  e2e           : same metric through the NODE: SDMatteApply.apply_matte(..., mask_refine=True) with pageable ComfyUI-style host
                  tensors in and host tensors out (staging + H2D + forward + post-processing + D2H inside the timed region).
  gpu_baseline  : the real competitor: the reference graph under torch.autocast(fp16) with
                  SlicedAttnProcessor(1) semantics (oracle mode "autocast": cuDNN / cuBLAS) on the SAME B200, CUDA events.
  cpu_baseline  : the oracle (torch fp32 restatement of the reference graph, kind "port") timed on the host cores.
  worst_case    : the same step with an all-foreground trimap (attn1 streams 100 % of the keys instead of ~30 %).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TFLOP_PER_MATTE = {128: 0.3477, 256: 1.4102, 384: 3.2458, 512: 5.952, 640: 9.664, 768: 14.558, 896: 20.848, 1024: 28.785}  # SURVEY.md §8(d) / App. D model


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_sustained": d.get("bf16_tflops_sustained", 1438.8), "tflops_burst": d.get("bf16_tflops", 1677.1),
                "hbm_gbs": d.get("hbm_gbs", 6566.1), "source": "measured"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def cpu_time_ratio_1024_over_512():
    """Measured wall-time ratio of one 1024^2 (sliced attention) to one 512^2 matte of the CPU oracle on a gpurun host
    (profiles/r2_cpu_1024.json, written by profiles/scripts/run_r2_cpu1024.sh); None if that measurement is not in the tree."""
    p = os.path.join(ROOT, "profiles", "r2_cpu_1024.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["seconds_per_matte_1024"] / d["seconds_per_matte_512"], p
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def shard_range(total: int, rank: int, world: int):
    """Batch shard of rank `rank`: samples [lo, hi). Samples are independent, so this is the whole partitioning."""
    per = total // world
    assert per * world == total, "global batch must divide by the number of GPUs"
    return rank * per, (rank + 1) * per


def cpu_oracle_rate(R: int, n_mattes: int, threads: int, sd=None, min_seconds: float = 0.0, max_mattes: int = 16, max_seconds: float = 1e9):
    """Time the oracle (fp32; sliced attention above 512^2, as the fp32 score tensor of one un-sliced L0 attention would be 5.4 GB)
    on the host: at least one matte, at most `n_mattes`; stops early once `max_seconds` of wall time are spent, continues beyond
    `n_mattes` (up to `max_mattes`) until `min_seconds` are filled.  Returns (mattes/s, seconds, mattes done, description)."""
    import torch
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    torch.set_num_threads(threads)
    if sd is None:
        sd = synth.make_checkpoint(seed=1234)
    image, trimap = synth.make_inputs(1, R, seed=0)
    t0 = time.perf_counter()
    done = 0
    while True:
        orc.forward(sd, image, trimap, is_transparent=False, sliced=R > 512)
        done += 1
        el = time.perf_counter() - t0
        if el >= max_seconds:
            break
        if done >= n_mattes and (el >= min_seconds or done >= max_mattes):
            break
    dt = time.perf_counter() - t0
    return done / dt, dt, done, f"{done} matte(s) at {R}x{R}, fp32 torch CPU oracle, {threads} threads, {dt:.1f} s"


def run_reference(args):
    """--impl reference: the reference's own CPU path does not run as written (hard-coded .cuda(), meta_arch.py:128; diffusers is
    absent), so this times the oracle port of the same graph on the host cores, on THE HEADLINE GEOMETRY: every step is one real
    1024 x 1024 matte (one sample of the bs=8 batch; samples are independent, so a batch costs 8x) with per-head sliced attention.
    One such matte takes minutes on 16-32 host cores, so the run is time-bounded: mattes are executed until `--ref-budget-s`
    (default 240 s) is spent, at least one; `steps_executed` says how many of the K requested steps really ran and nothing is
    extrapolated from another resolution."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import synth

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synth.make_checkpoint(seed=1234)
    R = args.ref_size or args.size
    rate, dt, done, desc = cpu_oracle_rate(R, max(1, args.steps), threads, sd, max_seconds=args.ref_budget_s)
    where = "at the workload's own resolution" if R == args.size else f"at {R}x{R} INSTEAD of the workload's {args.size}x{args.size} (--ref-size; not comparable)"
    sample = (f"{desc}; each step = ONE matte of the bs={args.batch} batch {where}; {done} of {args.steps} requested "
              f"steps executed (time-bounded at {args.ref_budget_s:.0f} s, no warm-up: a matte takes minutes)")
    line = {
        "impl": "reference", "metric": f"mattes/sec @{args.size}^2 bs={args.batch}", "value": rate, "unit": "mattes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_executed": done, "ms_per_step": 1000.0 * dt / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"bs={args.batch} {args.size}x{args.size} synthetic RGB+trimap, synthetic SDMatte checkpoint (seed 1234), is_transparent=False",
                   "resolution": R},
        "cpu_baseline": {"value": rate, "unit": "mattes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "mattes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def gpu_baseline(R: int, batches, dev, steps: int = 2):
    """The competitor on the same GPU: the reference graph under a genuine torch.autocast(fp16) with per-(sample, head) sliced
    attention (oracle mode "autocast" == sdmatte_nodes.py:331-358), fp32 master weights, cuDNN / cuBLAS kernels, CUDA events."""
    import torch
    from oracle import sdmatte_oracle as orc
    from oracle import synth

    sd = orc.to_device(synth.make_checkpoint(seed=1234), dev)
    out = {"impl": "oracle graph, torch.autocast(cuda, fp16) + SlicedAttnProcessor(1) semantics, cuDNN/cuBLAS, inputs resident on the device",
           "torch": torch.__version__, "resolution": R}
    for B in batches:
        image, trimap = synth.make_inputs(B, R, seed=1000)
        img, tri = image.to(dev), trimap.to(dev)
        orc.forward(sd, img, tri, mode="autocast", device=dev)  # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            orc.forward(sd, img, tri, mode="autocast", device=dev)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        out[f"bs{B}"] = {"ms_per_step": ms, "mattes_per_s": B / (ms * 1e-3), "steps": steps, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
        del img, tri
        torch.cuda.empty_cache()
    del sd
    torch.cuda.empty_cache()
    return out


def run_sweep(args, pkg, nodes, dev, world, rank, local_rank):
    """BASELINE config 5: the same engine over a list of inference sizes (bs per GPU fixed), device-timed like the headline run,
    one JSON line with a `sweep` list (value, ms_per_step, path roofline per resolution).  No e2e / baselines here."""
    import torch
    import torch.distributed as dist
    from oracle import synth

    B = args.batch
    nodes.register_state_dict("SDMatte.safetensors", synth.make_checkpoint(seed=1234))
    eng = nodes.get_engine("SDMatte.safetensors", dev)
    peaks = measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for R in [int(x) for x in args.sweep.split(",")]:
        lo, _ = shard_range(world * B, rank, world)
        image, trimap = synth.make_inputs(B, R, seed=1000 + lo)
        img_d, tri_d = image.to(dev), trimap.to(dev)
        alpha_d = torch.empty((B, R, R), dtype=torch.float16, device=dev)
        gathered = torch.empty((world * B, R, R), dtype=torch.float16, device=dev) if world > 1 else None

        def step():
            flush.zero_()
            eng.forward(img_d, tri_d, False, out=alpha_d)
            if world > 1:
                dist.all_gather_into_tensor(gathered, alpha_d)

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = t.item() / args.steps
        tf = world * B * TFLOP_PER_MATTE[R] / (ms_step * 1e-3)
        out.append({"resolution": R, "value": world * B / (ms_step * 1e-3), "unit": "mattes/s", "ms_per_step": ms_step, "global_batch": world * B,
                    "algorithmic_tflop_per_matte": TFLOP_PER_MATTE[R], "achieved_tflops_all_gpus": tf,
                    "frac_of_sustained_peak": tf / (world * peaks["tflops_sustained"]), "launches_per_step": eng.stats()["launches"]})
        del img_d, tri_d, alpha_d, gathered
        eng._ws = None
        torch.cuda.empty_cache()
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        print(json.dumps({"metric": f"mattes/sec resolution sweep bs={B} per GPU", "unit": "mattes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "higher_is_better": True, "scaling": "weak", "dtype": "f16", "data": "synthetic", "clocks": clocks,
                          "config": {"workload": f"bs={B} per GPU, synthetic RGB+trimap at each resolution, synthetic SDMatte checkpoint (seed 1234)",
                                     "parallelism": f"dp{world} batch shard + one NCCL all-gather of alpha" if world > 1 else "single GPU",
                                     "l2": "256 MiB buffer written between timed iterations"},
                          "peak": peaks, "sweep": out}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    from oracle import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pkg = ge.load_package()
    nodes = pkg.sdmatte_nodes

    R, B = args.size, args.batch
    if args.sweep:
        return run_sweep(args, pkg, nodes, dev, world, rank, local_rank)
    sd = synth.make_checkpoint(seed=1234)
    nodes.register_state_dict("SDMatte.safetensors", sd)
    nodes.set_devices([dev])
    eng = nodes.get_engine("SDMatte.safetensors", dev)  # the same cached engine the node call uses
    if not args.graph:
        eng.set_option("cuda_graph", 0)
    del sd
    # global batch = world*B; this rank's shard (weak scaling: B per GPU)
    lo, hi = shard_range(world * B, rank, world)
    image, trimap = synth.make_inputs(B, R, seed=1000 + lo)
    if args.trimap == "allfg":
        trimap = torch.ones_like(trimap)
    # attn1 only streams the keys whose probability can be non-zero under the -10000 trimap bias (key_compact_kernel):
    # fraction of the level-0 keys kept for this batch (same rule on the host: mask >= max - 0.25), reported in `config`
    def kept_fraction(t):
        m0 = t[:, ::8, ::8].reshape(t.shape[0], -1)
        return float((m0 >= m0.max(dim=1, keepdim=True).values - 0.25).float().mean())

    kept_l0 = kept_fraction(trimap)
    img_d, tri_d = image.to(dev), trimap.to(dev)
    alpha_d = torch.empty((B, R, R), dtype=torch.float16, device=dev)
    gathered = torch.empty((world * B, R, R), dtype=torch.float16, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)

    def step_device(tri=tri_d):
        flush.zero_()  # L2 flush between iterations (inside the loop; ~0.1 ms of a >100 ms step)
        eng.forward(img_d, tri, False, out=alpha_d)
        if world > 1:
            dist.all_gather_into_tensor(gathered, alpha_d)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    stats = eng.stats()
    graph = eng.graph_stats()

    # ---- worst case for the key compaction: all-foreground trimap (every key kept at every level), same step otherwise
    worst = None
    if args.trimap != "allfg" and not args.quick and not args.no_worst_case:
        tri_fg = torch.ones_like(tri_d)
        for _ in range(2):
            step_device(tri_fg)
        wsteps = max(2, min(5, args.steps))
        ms_w = timed(lambda: step_device(tri_fg), wsteps)
        worst = {"trimap": "all foreground (attn1 streams 100 % of the keys)", "ms_per_step": ms_w / wsteps,
                 "value": world * B * wsteps / (ms_w * 1e-3), "unit": "mattes/s", "steps": wsteps}
        del tri_fg

    # ---- end to end through the NODE (the call a ComfyUI user makes): pageable host tensors in, host tensors out
    node = nodes.SDMatteApply()
    img_h, tri_h = image.clone(), trimap.clone()  # plain pageable tensors, like ComfyUI's
    e2e_alpha = [None]

    def step_node():
        a, _ = node.apply_matte("SDMatte.safetensors", img_h, tri_h, R, False, "alpha_only", True, 0.8)
        e2e_alpha[0] = a
        if world > 1:
            dist.all_gather_into_tensor(gathered, a.to(dev, non_blocking=True))

    e2e_steps = 1 if args.quick else args.steps
    for _ in range(0 if args.quick else min(2, args.warmup)):
        step_node()
    ms_e2e = timed(step_node, e2e_steps)
    node_split = eng.node_call_timing()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    # ---- per-kernel-family breakdown from one profiled step (CUDA events around every launch)
    prof = eng.forward_profiled(img_d, tri_d, False)
    if args.dump_ops:
        with open(args.dump_ops, "w") as f:
            f.write("idx,kind,ms,gflop,mbytes,tflops,gbs\n")
            for i, (kind, kms, fl, by) in enumerate(prof):
                f.write(f"{i},{kind},{kms:.4f},{fl / 1e9:.2f},{by / 1e6:.2f},{fl / (kms * 1e-3) / 1e12 if kms > 0 else 0:.1f},{by / (kms * 1e-3) / 1e9 if kms > 0 else 0:.1f}\n")
    fam = {}
    for kind, kms, fl, by in prof:
        f = fam.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        f["ms"] += kms; f["flops"] += fl; f["bytes"] += by; f["launches"] += 1
    tot_ms = sum(f["ms"] for f in fam.values())
    breakdown = {}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(f["ms"], 3), "share": round(f["ms"] / tot_ms, 4), "launches": f["launches"]}
        if f["flops"] > 0:
            e["tflops"] = round(f["flops"] / (f["ms"] * 1e-3) / 1e12, 1)
        elif f["bytes"] > 0:
            e["gbs"] = round(f["bytes"] / (f["ms"] * 1e-3) / 1e9, 1)
        breakdown[k] = e

    def family(pred):
        ms = sum(f["ms"] for k, f in fam.items() if pred(k))
        fl = sum(f["flops"] for k, f in fam.items() if pred(k))
        n = sum(f["launches"] for k, f in fam.items() if pred(k))
        return ms, fl, n

    # the 3x3 stride-1 convolution kernels, incl. the polyphase launches of the upsamplers (same kernel, 4 taps; EXECUTED flops)
    conv_ms, conv_fl, conv_n = family(lambda k: k in ("tc:conv3x3", "tc:conv3x3_poly"))
    executed_tflop = sum(f["flops"] for f in fam.values()) / 1e12
    gemm_ms, gemm_fl, gemm_n = family(lambda k: k.startswith("tc:") and "attention" not in k)
    conv_ach = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    gemm_ach = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tp) and R == 1024 and B == 8:
        try:
            td = json.load(open(tp))
            # ncu recognises the family by kernel name (conv_swap*_kernel + the HALO instantiations): the 65 big launches of the 98;
            # the other 33 (small geometries on the tap-per-box kernel) carry 1.6 % of the family's algorithmic bytes
            traffic = {"dram_bytes_per_launch": td["conv3x3"]["dram_bytes"] / td["conv3x3"]["launches"], "launches_measured": td["conv3x3"]["launches"],
                       "dram_bytes_per_step_measured_launches": td["conv3x3"]["dram_bytes"],
                       "algorithmic_bytes_per_step_all_launches": fam["tc:conv3x3"]["bytes"], "launches_per_step": conv_n,
                       "measured_at": "r2l: before the CTA-pair form and the polyphase upsamplers (which removed three 3x3 launches over upsampled tensors)",
                       "source": "profiles/r2_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one bs=8 1024^2 forward, "
                                 "profiles/scripts/run_r2_traffic.sh): measured traffic is BELOW the read-once/write-once figure (producer outputs still in the 126 MB L2)"}
        except Exception:
            traffic = None
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)
    path_tflops = B * TFLOP_PER_MATTE.get(R, 0.0) / (ms_step * 1e-3)

    # ---- the real competitor on the same GPU
    gpu_base = None
    if world == 1 and not args.no_gpu_baseline and not args.quick:
        eng._ws = None  # hand the workspace back to the allocator while the baseline runs
        torch.cuda.empty_cache()
        try:
            gpu_base = gpu_baseline(R, sorted({1, B}), dev)
            gpu_base["speedup_device_timed"] = value / gpu_base[f"bs{B}"]["mattes_per_s"]
            gpu_base["speedup_e2e_node_vs_device_resident_baseline"] = e2e_value / gpu_base[f"bs{B}"]["mattes_per_s"]
        except Exception as ex:  # the baseline must never take the bench line down
            gpu_base = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- CPU baseline: the oracle on the host cores, bounded sample (N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        threads = os.cpu_count() or 1
        Rc = 512
        rate, dt, done, desc = cpu_oracle_rate(Rc, 1, threads, min_seconds=10.0, max_mattes=8)  # about 10-30 s of CPU work
        ratio, src = cpu_time_ratio_1024_over_512()
        if ratio is not None and R == 1024:
            equiv, how = rate / ratio, f"converted to {R}^2 mattes with the MEASURED wall-time ratio t(1024^2)/t(512^2) = {ratio:.1f} of the same oracle ({os.path.relpath(src, ROOT)}); `bench.py --impl reference` times real 1024^2 mattes"
        else:
            equiv, how = rate * TFLOP_PER_MATTE[Rc] / TFLOP_PER_MATTE[R], f"converted to {R}^2 mattes by the algorithmic FLOP ratio {TFLOP_PER_MATTE[Rc]}/{TFLOP_PER_MATTE[R]} (an upper bound: sliced attention at 1024^2 is slower than that)"
        cpu = {"value": equiv, "unit": "mattes/s", "cores": threads, "kind": "port", "sample": desc + "; " + how,
               "measured_512": {"value": rate, "unit": "mattes/s at 512^2"}}

    line = {
        "metric": f"mattes/sec @{R}^2 bs={B}", "value": value, "unit": "mattes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"bs={B} per GPU, {R}x{R} synthetic RGB+trimap, synthetic SDMatte checkpoint (seed 1234), is_transparent=False, mask_refine on (e2e leg)",
                   "global_batch": world * B, "resolution": R, "parallelism": f"dp{world} batch shard + one NCCL all-gather of alpha" if world > 1 else "single GPU",
                   "l2": "256 MiB buffer written between timed iterations (inside the loop)",
                   "trimap": args.trimap, "self_attn_keys_streamed_L0": round(kept_l0, 4),
                   "cuda_graph": {"enabled": bool(args.graph), **graph},
                   "note": "FLOP figures are ALGORITHMIC (all keys); attn1 skips keys whose softmax probability is exactly 0 under the "
                           "reference's -10000 bias, so tc:attention_self reports algorithmic TFLOP/s above what it executes; see worst_case"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "mattes/s", "h2d_bytes_per_step": B * R * R * 16, "d2h_bytes_per_step": B * R * R * 2,
                "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "host_split_last_call_ms": node_split,
                "api": "SDMatteApply.apply_matte(ckpt, image, trimap, R, False, 'alpha_only', mask_refine=True, 0.8) with pageable host tensors"},
        "gpu_launches": stats["launches"] * args.steps,
        "roofline": {"bound": "tensor", "kernel": "tc:conv3x3 — the tcgen05 implicit-GEMM 3x3 stride-1 convolutions (conv_gemm_kernel<..., HALO>, conv_swap_kernel), all launches of one step",
                     "achieved": conv_ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": conv_ach / peaks["tflops_sustained"],
                     "traffic": traffic, "launches_per_step": conv_n, "share_of_step": conv_ms / tot_ms if tot_ms else None,
                     "peak_source": peaks["source"] + " (bf16 sustained: the kernels are timed inside a long step)"},
        "gemm_roofline": {"kernel": "all non-attention tcgen05 kernels (3x3 / 1x1 convs, linears, GEGLU, V^T, VAE QK^T / PV)", "achieved": gemm_ach,
                          "frac": gemm_ach / peaks["tflops_sustained"], "launches_per_step": gemm_n, "share_of_step": gemm_ms / tot_ms if tot_ms else None},
        "path_roofline": {"algorithmic_tflop_per_matte": TFLOP_PER_MATTE.get(R), "executed_tflop_per_matte": round(executed_tflop / B, 3),
                          "note": "algorithmic = the reference's formulation (SURVEY 8d); the engine executes fewer tensor FLOPs: attn1 streams only the keys "
                                  "that can have non-zero probability, the upsamplers run as four 2x2-tap polyphase convs (4/9 of the 3x3 form)",
                          "achieved_tflops": path_tflops,
                          "frac_of_sustained_peak": path_tflops / peaks["tflops_sustained"]},
        "kernel_breakdown": breakdown,
        "worst_case": worst,
        "gpu_baseline": gpu_base,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=8, help="mattes per GPU per step")
    ap.add_argument("--trimap", default="synthetic", choices=["synthetic", "allfg"], help="allfg: every pixel foreground (worst case of the key compaction)")
    ap.add_argument("--graph", type=int, default=1, help="0: launch the plan kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-worst-case", action="store_true")
    ap.add_argument("--dump-ops", default=None, help="write the per-op profile of one step to this CSV")
    ap.add_argument("--ref-size", type=int, default=0, choices=[0, 128, 256, 384, 512, 1024], help="--impl reference: resolution of the timed mattes (0 = --size)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: wall-time bound of the whole run")
    ap.add_argument("--sweep", default="", help="comma-separated inference sizes (BASELINE config 5), e.g. 512,640,768,896,1024: device-timed only")
    ap.add_argument("--quick", action="store_true", help="profiling aid: no e2e repeat / baselines, warm-up as given (not a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3 and not args.quick:
        args.warmup = 3
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
