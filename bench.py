#!/usr/bin/env python
"""bench.py -- photometric-loss fwd+bwd throughput of the CoDEPS hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic frame triplets: the
reconstruction loss (5 scales, 2 source frames, auto-mask) + the edge-aware smoothness loss,
forward and backward (loss values, per-level argmin masks, dL/d depth, dL/d disp, dL/dT).
The default workload is BASELINE.json configs[1]: Cityscapes-shaped 1024x512 triplets, batch 8
per GPU (weak scaling: every rank gets its own batch; no data-path collective).

Output: ONE JSON line on rank 0 (see the keys at the bottom of main()).
  value     triplets/s over all GPUs, inputs resident in HBM, the step replayed as a CUDA graph
            captured from the public classes; device time from CUDA events, max over ranks.
  e2e       same metric through the public classes with HOST (pinned) inputs: per-step H2D copy
            of all inputs and D2H read of the loss inside the timed region.
  roofline  the dominant kernel (fused tile kernel): algorithmic bytes per launch / its mean
            duration from CUDA events recorded around every launch in an eager pass of the same
            K steps; peak from MEASURED_PEAKS.json.
  cpu_baseline  the oracle port of the reference torch path timed on the host cores (rank 0, N=1).
  extras    short runs of the other configurations BASELINE.json names (workloads below), the
            reference algorithm in torch eager on cuda:0 (informational) and the full adaptation
            step (bench_adapt.py); `--workload NAME` runs one of them as the main line instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

# name: groups of (preset, batch, flip every other cx, key), batch mode, description.
#   "per_gpu": every rank processes `batch` triplets of every group (weak scaling);
#   "global":  `batch` is the global batch, split contiguously over the ranks (strong scaling).
WORKLOADS = {
    "cityscapes_b8": dict(groups=[("cityscapes", 8, False, "train")], mode="per_gpu",
                          desc="Cityscapes-shaped 1024x512 triplets, batch 8 per GPU, 5 scales (BASELINE config 2)"),
    "kitti360_b8": dict(groups=[("kitti360", 8, True, "train")], mode="per_gpu",
                        desc="KITTI-360-shaped 1408x376 triplets, batch 8 per GPU, per-sample intrinsics"),
    "semkitti_b8": dict(groups=[("semkitti", 8, True, "train")], mode="per_gpu",
                        desc="SemKITTI-DVPS-shaped 1280x384 triplets, batch 8 per GPU"),
    "kitti360_b16_mixed": dict(groups=[("kitti360", 16, True, "online+replay")], mode="global",
                               desc="KITTI-360-shaped 1408x376 triplets, GLOBAL batch 16 (online + replay samples, every "
                                    "other one with the flipped principal point) sharded over the GPUs (BASELINE config 3)"),
    "semkitti_b64": dict(groups=[("semkitti", 64, True, "train")], mode="global",
                         desc="SemKITTI-DVPS-shaped 1280x384 triplets, GLOBAL batch 64 sharded over the GPUs, loss only "
                              "(BASELINE config 5)"),
    "adapt_mix": dict(groups=[("cityscapes", 2, False, "source"), ("kitti360_cfg", 1, False, "target"),
                              ("kitti360_cfg", 2, True, "target_replay")], mode="per_gpu",
                      desc="online-adaptation batch per GPU: 2 source @1024x512 + 1 target + 2 target-replay @1408x384, "
                           "losses combined as sum n_k L_k / sum n_k (algos/depth.py:562-568)"),
}
EXTRA_WORKLOADS = ("kitti360_b16_mixed", "semkitti_b64", "adapt_mix")
RECON_WEIGHT, SMOOTH_WEIGHT = 10.0, 0.001  # cfg/train_cityscapes.yaml:40-41
NUM_SCALES = 5
INPUT_SETS = 3  # rotating input sets so that consecutive steps do not hit L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["adapt_step"], default="cityscapes_b8")
    ap.add_argument("--noise", choices=("torch", "fused"), default="fused",
                    help="tie-break noise of the identity candidates: the library default (in-kernel generator) or "
                         "torch.randn per level as the reference draws it")
    ap.add_argument("--intrinsics", choices=("host", "device"), default="host",
                    help="host: CameraModel objects hold host values (kernel parameter space); device: lazy "
                         "CameraModel.from_tensor of CUDA rows, the kernels read the calibration from HBM")
    ap.add_argument("--overlap-smooth", action="store_true",
                    help="run the smoothness loss on a second stream next to the reconstruction loss")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other configurations")
    ap.add_argument("--shift-px", type=int, default=3,
                    help="pixel shift between the synthetic frames (the pose explains it). The tile kernel serves "
                         "bilinear footprints within 4-6 px of the pixel from shared memory and the rest from global memory")
    return ap.parse_args()


def profiled_kernel(workload: str, kernel_substr: str):
    """Latest tracked ncu record (profiles/*_ncu.json, written by tools/ncu_to_profile.py) of a
    kernel for this workload: DRAM traffic per launch, issue-slot utilisation, L2 hit rate."""
    import glob
    import re
    best = None

    def natural(path):  # r01_9_... before r01_11_...
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(path))]
    blobs = []
    for path in glob.glob(os.path.join(REPO, "profiles", "*_ncu.json")):
        try:
            with open(path) as f:
                blobs.append((path, json.load(f)))
        except (OSError, ValueError):
            continue
    # newest record last: by its "created" stamp (records written before the stamp existed sort first), then by name
    for path, blob in sorted(blobs, key=lambda pb: (pb[1].get("created", ""), natural(pb[0]))):
        if blob.get("workload") != workload:
            continue
        for rec in blob.get("launches", []):
            if kernel_substr in rec.get("kernel", "") and "dram_bytes" in rec:
                best = dict(rec, profile=os.path.relpath(path, REPO))
    return best


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int, period_s: float = 0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period_s
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "samples": len(self.samples), "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index: int) -> dict:
    """Pin this process to the CPUs NVML reports as local to GPU `index` (so that pinned host buffers
    are first-touched on that NUMA node) and report the topology.  Best effort."""
    info = {"cpu_affinity_set": False, "gpu_numa_node": None}
    try:
        info["_original"] = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node_file = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
        if os.path.exists(node_file):
            info["gpu_numa_node"] = int(open(node_file).read().strip())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        info["cpu_affinity_set"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as exc:
        info["error"] = repr(exc)[:120]
    return info


def local_groups(workload: str, rank: int, world: int):
    """[(preset, local batch, flip, key, global offset)] of this rank for a workload."""
    from codeps_b200.distributed import shard_bounds
    spec = WORKLOADS[workload]
    out = []
    for preset, batch, flip, key in spec["groups"]:
        if spec["mode"] == "global":
            lo, hi = shard_bounds(batch, rank, world)
            if hi > lo:
                out.append((preset, hi - lo, flip, key, lo))
        else:
            out.append((preset, batch, flip, key, 0))
    return out


def global_triplets(workload: str, world: int) -> int:
    spec = WORKLOADS[workload]
    total = sum(g[1] for g in spec["groups"])
    return total if spec["mode"] == "global" else total * world


SHIFT_PX = {"value": 3}  # pixel shift between the synthetic frames (--shift-px)


def make_group_batch(preset, batch, flip, seed, offset=0):
    """`batch` triplets of a preset; with `flip`, samples with odd GLOBAL index get the mirrored
    principal point (replay samples, datasets/preprocessing.py:47-52)."""
    from codeps_b200 import synthetic
    tb = synthetic.make_preset_batch(preset, batch, seed=seed, shift_px=SHIFT_PX["value"])
    if flip:
        k = tb.intrinsics.clone()
        odd = (torch.arange(batch) + offset) % 2 == 1
        k[odd, 2] = tb.width - k[odd, 2] - 1
        tb = synthetic.TripletBatch(tb.images, tb.disp, tb.depth, tb.poses, k, tb.width, tb.height)
    return tb


# ------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle port of the reference torch path on the host cores
# ------------------------------------------------------------------------------------------
def reference_step_fn(workload: str, device: str, rank: int = 0, world: int = 1, per_group_batch=None):
    """step() -> (recon, smooth) running the oracle port (the reference's ATen op sequence) fwd+bwd
    on this rank's batch of the workload on `device`; returns (step, triplets per step)."""
    from oracle import photo_oracle as po
    groups = []
    for preset, batch, flip, key, off in local_groups(workload, rank, world):
        if per_group_batch is not None:
            batch = min(batch, per_group_batch)
        tb = make_group_batch(preset, batch, flip, seed=1000 * rank, offset=off)
        noise = po.draw_noise(batch, tb.width, tb.height, NUM_SCALES, seed=1)
        mv = lambda t: t.to(device)
        groups.append(dict(n=batch, k=tb.intrinsics.numpy(), images=[mv(i) for i in tb.images], depth=mv(tb.depth),
                           disp=mv(tb.disp), poses=[mv(p) for p in tb.poses], noise=[mv(n) for n in noise]))
    total = sum(g["n"] for g in groups)

    def step():
        recon_sum, smooth_sum = 0.0, 0.0
        for g in groups:
            share = g["n"] / total  # sum n_k L_k / sum n_k
            r = po.loss_and_grads(g["k"], g["images"], g["depth"], g["disp"], g["poses"], g["noise"], NUM_SCALES,
                                  recon_weight=RECON_WEIGHT * share, smooth_weight=SMOOTH_WEIGHT * share)
            recon_sum += share * float(r["recon"])
            smooth_sum += share * float(r["smooth"])
        return recon_sum, smooth_sum
    return step, total


def cpu_reference_rate(workload: str, steps: int, warmup: int, budget_s: float, per_group_batch=None):
    """Times the reference algorithm (oracle port, same ATen ops as the reference) fwd+bwd on the host
    cores, all threads, on this workload's per-GPU batch (or `per_group_batch` triplets of every
    group for a bounded sample).  Returns (triplets/s, cores, iterations, median s, triplets/step)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, total = reference_step_fn(workload, "cpu", per_group_batch=per_group_batch)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if i >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    med = statistics.median(times)
    return total / med, cores, len(times), med, total


def run_reference_arm(args):
    """The reference's own implementation of the path on the box's host cores: same workload, same
    per-GPU batch per step (same_config), bounded by a time budget."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from codeps_b200 import synthetic
    workload = args.workload if args.workload != "adapt_step" else "adapt_mix"
    spec = WORKLOADS[workload]
    preset = spec["groups"][0][0]
    w, h = synthetic.PRESETS[preset][0], synthetic.PRESETS[preset][1]
    # per-GPU batch of the N=1 run of our arm (global workloads are sharded there, not here: one host)
    rate, cores, iters, med, total = cpu_reference_rate(workload, args.steps, max(args.warmup, 1), budget_s=150.0)
    sample = (f"the full per-GPU batch ({total} triplets) per step, {iters} steps timed within the time budget, "
              f"oracle port of the reference torch-CPU path (oracle/photo_oracle.py), torch {torch.__version__}")
    line = {
        "impl": "reference", "metric": f"photometric-loss fwd+bwd frame-triplets/sec @{w}x{h}", "value": rate,
        "unit": "triplets/s", "n_gpus": args.gpus, "steps": iters, "warmup": max(args.warmup, 1),
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "description": spec["desc"], "per_gpu_batch": total, "num_scales": NUM_SCALES,
                   "device": "host CPU", "torch_threads": cores, "same_config": True},
        "cpu_baseline": {"value": rate, "unit": "triplets/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class LossWorkload:
    """Device-resident inputs and loss objects of one workload on this rank."""

    def __init__(self, name, rank, world, dev, noise="fused", intrinsics="host", input_sets=INPUT_SETS, pin=True):
        import codeps_b200
        from codeps_b200 import synthetic
        self.name, self.dev = name, dev
        self.groups = local_groups(name, rank, world)
        self.triplets = sum(g[1] for g in self.groups)
        self.input_sets = input_sets
        self.host, self.devs, self.cams, self.fns = [], [], [], []
        for gi, (preset, batch, flip, key, off) in enumerate(self.groups):
            w, h = synthetic.PRESETS[preset][0], synthetic.PRESETS[preset][1]
            hs = [make_group_batch(preset, batch, flip, seed=1000 * rank + 10 * gi + i, offset=off) for i in range(input_sets)]
            if pin:
                hs = [x.pin() for x in hs]
            ds = [x.to(dev) for x in hs]
            if intrinsics == "device":
                cams = [[codeps_b200.CameraModel.from_tensor(w, h, k[i]) for i in range(batch)]
                        for k in (x.intrinsics.to(dev) for x in hs)]
            else:
                cams = [x.camera_models() for x in hs]
            self.host.append(hs)
            self.devs.append(ds)
            self.cams.append(cams)
            self.fns.append(codeps_b200.ReconstructionLoss(w, h, codeps_b200.SSIMLoss(), NUM_SCALES, dev, noise=noise))
        self.smooth_fn = codeps_b200.EdgeAwareSmoothnessLoss()
        self.w_recon = torch.tensor(RECON_WEIGHT, device=dev)
        self.w_smooth = torch.tensor(SMOOTH_WEIGHT, device=dev)
        self.resident_bytes = sum(x.nbytes() for hs in self.host for x in hs)
        self.h2d_bytes = sum(hs[0].nbytes() for hs in self.host)

    def with_noise(self, noise):
        """Shallow copy sharing the resident inputs, with loss objects in another noise mode."""
        import copy
        import codeps_b200
        other = copy.copy(self)
        other.fns = [codeps_b200.ReconstructionLoss(f.scaled_width[0], f.scaled_height[0], codeps_b200.SSIMLoss(),
                                                    NUM_SCALES, self.dev, noise=noise) for f in self.fns]
        return other

    def step(self, i, side_stream=None):
        """fwd + bwd on resident input set i; returns (recon, smooth, grads).  Several groups are
        combined as DepthAlgo.adaptation does: sum n_k L_k / sum n_k (algos/depth.py:562-568)."""
        i %= self.input_sets
        leaves, recons, smooths = [], [], []
        for gi, (preset, batch, flip, key, off) in enumerate(self.groups):
            ds = self.devs[gi][i]
            # fresh autograd leaves every step (views, no copies), created on the launching stream
            depth, disp = ds.depth.detach().requires_grad_(True), ds.disp.detach().requires_grad_(True)
            p0, p1 = ds.poses[0].detach().requires_grad_(True), ds.poses[1].detach().requires_grad_(True)
            if side_stream is not None:
                # the two losses are independent: the smoothness kernels run on a second stream and
                # fill the SMs the tile kernel's last wave and the small reduction kernels leave idle
                cur = torch.cuda.current_stream()
                side_stream.wait_stream(cur)
                with torch.cuda.stream(side_stream):
                    smooth = self.smooth_fn(ds.images[0], disp)
                recon = self.fns[gi](self.cams[gi][i], ds.images, depth, (p0, p1))
                cur.wait_stream(side_stream)
            else:
                recon = self.fns[gi](self.cams[gi][i], ds.images, depth, (p0, p1))
                smooth = self.smooth_fn(ds.images[0], disp)
            leaves += [depth, disp, p0, p1]
            recons.append(recon * (batch / self.triplets) if len(self.groups) > 1 else recon)
            smooths.append(smooth * (batch / self.triplets) if len(self.groups) > 1 else smooth)
        recon = torch.stack(recons).sum() if len(recons) > 1 else recons[0]
        smooth = torch.stack(smooths).sum() if len(smooths) > 1 else smooths[0]
        # the caller's  loss = 10*recon + 0.001*smooth; loss.backward()  (train_codeps.py:102-107)
        grads = torch.autograd.grad([recon, smooth], leaves, grad_outputs=[self.w_recon, self.w_smooth])
        return recon, smooth, grads


def make_runner(step_fn, use_graph, input_sets):
    """Warm up, optionally capture one CUDA graph per input set, return run(i)."""
    for i in range(3):
        step_fn(i % input_sets)
    torch.cuda.synchronize()
    if not use_graph:
        return lambda i: step_fn(i % input_sets)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(input_sets):
            step_fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs, outs = [], []
    for i in range(input_sets):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs.append(step_fn(i))
        graphs.append(g)

    def run(i):
        graphs[i % input_sets].replay()
        return outs[i % input_sets]
    return run


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    import codeps_b200
    from codeps_b200 import _native, ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    _native.load()
    numa = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: first touch lands on the GPU's node

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(run, steps, warmup):
        """W warm-up steps, then exactly `steps` steps between barriers; device ms, max over ranks."""
        for i in range(max(warmup, 3)):
            run(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        last = None
        for i in range(steps):
            last = run(i)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)), last

    if args.workload == "adapt_step":
        if numa.get("_original"):
            os.sched_setaffinity(0, numa.pop("_original"))
        import bench_adapt
        line = bench_adapt.run(args, rank, world, dev, barrier, max_over_ranks)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    spec = WORKLOADS[args.workload]
    SHIFT_PX["value"] = args.shift_px
    use_graph = not args.no_graph
    wl = LossWorkload(args.workload, rank, world, dev, noise=args.noise, intrinsics=args.intrinsics)
    original_cpus = numa.pop("_original", None)
    if numa.get("cpu_affinity_set") and original_cpus:  # the pinned buffers exist: give the process all its cores back
        os.sched_setaffinity(0, original_cpus)
    preset = spec["groups"][0][0]
    w, h = synthetic.PRESETS[preset][0], synthetic.PRESETS[preset][1]
    total_triplets = global_triplets(args.workload, world)
    side_stream = torch.cuda.Stream() if args.overlap_smooth else None
    step = lambda i: wl.step(i, side_stream)

    torch.manual_seed(1234 + rank)
    for i in range(3):  # eager warm-up (sizes the allocator pools, sets kernel attributes)
        step(i)
    torch.cuda.synchronize()
    launches_before = ops.launch_count()
    step(0)
    launches_per_step = ops.launch_count() - launches_before
    torch.cuda.synchronize()
    run_step = make_runner(step, use_graph, wl.input_sets)

    # ---- timed region: device-resident inputs
    with ClockSampler(local_rank) as clocks:
        elapsed_ms, last = timed(run_step, args.steps, args.warmup)
    value = total_triplets * args.steps / (elapsed_ms * 1e-3)
    recon_val, smooth_val = float(last[0].detach()), float(last[1].detach())

    # ---- eager pass with per-kernel CUDA events (same K steps) -> dominant-kernel duration
    _native.profile_enable(True)
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    prof = _native.profile_read()
    _native.profile_enable(False)
    kernel_ms = {k: (ms / n if n else None) for k, (ms, n) in prof.items()}
    photo_ms = kernel_ms["photo"]  # mean per launch (one launch per group and step)
    peak, peak_src = measured_peaks()
    # fused tile kernel = the per-level forward (41 B/level-px) and backward (45 B/level-px) of
    # SURVEY.md section 8d done in one pass; per launch = one group
    alg_launch, a_alg_step = [], 0
    for g_preset, g_batch, _, _, _ in wl.groups:
        gw, gh = synthetic.PRESETS[g_preset][0], synthetic.PRESETS[g_preset][1]
        alg_launch.append(86 * sum(ww * hh for ww, hh in synthetic.level_sizes(gw, gh, NUM_SCALES)) * g_batch)
        a_alg_step += synthetic.algorithmic_bytes(gw, gh, NUM_SCALES) * g_batch
    photo_alg_bytes = sum(alg_launch) / len(alg_launch)
    achieved = photo_alg_bytes / (photo_ms * 1e-3) / 1e9
    step_kernel_ms = sum(v for v in kernel_ms.values() if v) * len(wl.groups)
    prof_rec = profiled_kernel(args.workload, "cdp_photo_kernel<1") or profiled_kernel(args.workload, "cdp_photo_kernel")
    local_rate = wl.triplets * args.steps / (elapsed_ms * 1e-3)
    roofline = {
        "bound": "hbm", "kernel": "cdp_photo_kernel<true,false>", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": prof_rec["dram_bytes"] if prof_rec else None, "peak_source": peak_src,
        "ncu": ({k: prof_rec.get(k) for k in ("profile", "duration_us", "issue_slot_pct", "sm_pct_of_peak",
                                              "dram_pct_of_peak", "l2_hit_pct", "l1_hit_pct",
                                              "achieved_occupancy_pct", "registers_per_thread", "warp_instructions")}
                if prof_rec else None),
        "limiter": "fp32 issue slots, not HBM (DESIGN.md section 6): see ncu.issue_slot_pct vs ncu.dram_pct_of_peak",
        "algorithmic_bytes_per_launch": photo_alg_bytes, "kernel_ms": photo_ms,
        "kernel_share_of_step": photo_ms * len(wl.groups) / step_kernel_ms if step_kernel_ms else None,
        "kernel_ms_all": kernel_ms,
        "path_bytes_per_step_per_gpu": a_alg_step,
        "path_frac": (a_alg_step / wl.triplets) * local_rate / 1e9 / peak,
    }

    extras = {}
    # ---- same steps with the tie-break noise drawn by torch.randn per level like the reference (noise="torch")
    if args.noise == "fused" and not args.no_extras:
        wl_torch = wl.with_noise("torch")  # same resident inputs
        torch_ms, _ = timed(make_runner(lambda i: wl_torch.step(i), use_graph, wl.input_sets), args.steps, args.warmup)
        extras["value_torch_noise"] = total_triplets * args.steps / (torch_ms * 1e-3)
        extras["note"] = ("value_torch_noise: same step with ReconstructionLoss(noise='torch') -- five torch.randn launches "
                          "per call (the reference's own random stream) and 16 B of noise traffic per level-pixel "
                          "instead of the in-kernel generator")
    elif not args.no_extras:
        wl_fused = wl.with_noise("fused")
        fused_ms, _ = timed(make_runner(lambda i: wl_fused.step(i), use_graph, wl.input_sets), args.steps, args.warmup)
        extras["value_fused_noise"] = total_triplets * args.steps / (fused_ms * 1e-3)

    # ---- end to end: host (pinned) inputs -> public classes -> loss read back on the host
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(wl, args, dev, barrier, max_over_ranks, total_triplets, world)
        e2e["numa"] = numa

    # ---- the other configurations BASELINE.json names, as short runs
    if not args.no_extras and args.workload == "cityscapes_b8":
        import gc

        def fresh():  # every extra starts from an empty caching allocator (they differ wildly in block sizes)
            gc.collect()
            torch.cuda.synchronize()
            torch.cuda.empty_cache()

        fresh()
        try:
            import bench_adapt
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup = 12, 4
            extras["adapt_step"] = bench_adapt.run(sub, rank, world, dev, barrier, max_over_ranks, brief=True)
        except Exception as exc:  # e.g. torchvision missing
            extras["adapt_step"] = {"unavailable": repr(exc)}
        fresh()
        extras["workloads"] = {}
        for name in EXTRA_WORKLOADS:
            extras["workloads"][name] = run_extra_workload(name, rank, world, dev, timed, use_graph)
        # sensitivity to the image motion: 12 px instead of 3 px between the frames puts every level-0
        # footprint outside the staged source boxes (global-memory taps), levels >= 2 stay inside
        SHIFT_PX["value"] = 12
        extras["workloads"]["cityscapes_b8_shift12px"] = run_extra_workload("cityscapes_b8", rank, world, dev, timed, use_graph)
        extras["workloads"]["cityscapes_b8_shift12px"]["description"] += "; frames shifted by 12 px instead of 3"
        SHIFT_PX["value"] = args.shift_px
        fresh()
        extras["torch_cuda_eager"] = run_torch_cuda_eager(args.workload, rank, world, dev, timed, value)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        rate, cores, iters, med, total = cpu_reference_rate(args.workload, steps=12, warmup=1, budget_s=20.0,
                                                            per_group_batch=2)
        cpu_baseline = {"value": rate, "unit": "triplets/s", "cores": cores, "kind": "port",
                        "sample": f"{total} triplet(s) of the workload per iteration, {iters} iterations, "
                                  f"median {med * 1e3:.0f} ms, torch {torch.__version__} CPU; the --impl reference arm "
                                  f"times the full per-GPU batch"}

    # scalar statistics only (one coalesced all-reduce of 3 doubles); the data path has no collective
    from codeps_b200.distributed import reduce_loss_dict
    stats = reduce_loss_dict({"recon": torch.tensor(recon_val, device=dev, dtype=torch.float64),
                              "smooth": torch.tensor(smooth_val, device=dev, dtype=torch.float64)}, max(wl.triplets, 1))
    checksum = [float(stats["recon"]), float(stats["smooth"])]

    if rank == 0:
        line = {
            "metric": f"photometric-loss fwd+bwd frame-triplets/sec @{w}x{h}",
            "value": value, "unit": "triplets/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if spec["mode"] == "global" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": spec["desc"], "per_gpu_batch": wl.triplets,
                       "global_batch": total_triplets, "num_scales": NUM_SCALES, "noise": args.noise,
                       "intrinsics": args.intrinsics, "overlap_smooth": bool(args.overlap_smooth),
                       "timed_with": "cuda_graph_replay" if use_graph else "eager_launches",
                       "l2": f"{wl.input_sets} rotating input sets, {wl.resident_bytes / 1e6:.0f} MB resident > 126 MB L2",
                       "loss_weights": [RECON_WEIGHT, SMOOTH_WEIGHT], "shift_px": args.shift_px},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "extras": extras,
            "loss": {"recon": float(checksum[0]), "smooth": float(checksum[1])},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(wl, args, dev, barrier, max_over_ranks, total_triplets, world):
    """Host (pinned) inputs -> public classes -> loss read back on the host, every step; the next
    step's upload runs on a copy stream while this step's kernels run."""
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    e2e_steps = max(10, min(args.steps, 40))
    # two device staging sets per group (ping-pong), allocated once: the upload of step i+1 overwrites
    # the buffers step i-1 computed on, after that step's kernels have finished (event)
    slots = [[hs[0].to(dev) for hs in wl.host] for _ in range(2)]
    free_evt = [None, None]

    def upload(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            if free_evt[slot] is not None:
                copy_stream.wait_event(free_evt[slot])
            for gi, hs in enumerate(wl.host):
                src, dst = hs[i % wl.input_sets], slots[slot][gi]
                for a, b in zip(list(dst.images) + [dst.disp, dst.depth] + list(dst.poses),
                                list(src.images) + [src.disp, src.depth] + list(src.poses)):
                    a.copy_(b, non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(copy_stream)
        return slot, evt

    loss_host = torch.zeros(2, pin_memory=True)

    def e2e_step(i, staged):
        slot, evt = staged
        nxt = upload(i + 1)  # overlap the next step's H2D with this step's kernels
        main_stream.wait_event(evt)
        recon_t, smooth_t = 0.0, 0.0
        for gi, dsb in enumerate(slots[slot]):
            share = wl.groups[gi][1] / wl.triplets
            depth = dsb.depth.detach().requires_grad_(True)
            disp = dsb.disp.detach().requires_grad_(True)
            p0, p1 = dsb.poses[0].detach().requires_grad_(True), dsb.poses[1].detach().requires_grad_(True)
            recon = wl.fns[gi](wl.cams[gi][i % wl.input_sets], dsb.images, depth, (p0, p1))
            smooth = wl.smooth_fn(dsb.images[0], disp)
            recon_t = recon_t + share * recon
            smooth_t = smooth_t + share * smooth
        (RECON_WEIGHT * recon_t + SMOOTH_WEIGHT * smooth_t).backward()
        loss_host.copy_(torch.stack((recon_t.detach(), smooth_t.detach())), non_blocking=True)
        free_evt[slot] = torch.cuda.Event()
        free_evt[slot].record(main_stream)
        return nxt

    # pinned H2D copy rate of this rank alone and of all ranks at once (the ceiling of the e2e number)
    probe = wl.host[0][0]
    rank = int(os.environ.get("RANK", "0"))

    def probe_copy():
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            for a, b in zip(list(slots[0][0].images) + [slots[0][0].disp, slots[0][0].depth],
                            list(probe.images) + [probe.disp, probe.depth]):
                a.copy_(b, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        return c0.elapsed_time(c1) / 5

    probe_copy()  # warm-up
    barrier()
    alone_ms = probe_copy() if rank == 0 else 0.0  # rank 0 copies while every other rank waits at the barrier
    barrier()
    all_ms = max_over_ranks(probe_copy())          # every rank copies at once
    barrier()
    probe_bytes = probe.nbytes() - sum(p_.numel() * 4 for p_ in probe.poses)
    h2d_alone = probe_bytes / (max_over_ranks(alone_ms) * 1e-3) / 1e9
    h2d_concurrent = probe_bytes / (all_ms * 1e-3) / 1e9

    staged = upload(0)
    for i in range(3):
        staged = e2e_step(i, staged)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(e2e_steps):
        staged = e2e_step(i, staged)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    n_global = total_triplets
    return {"value": n_global * e2e_steps / (e2e_ms * 1e-3), "unit": "triplets/s",
            "h2d_bytes_per_step": wl.h2d_bytes, "d2h_bytes_per_step": 8, "steps": e2e_steps,
            "ms_per_step": e2e_ms / e2e_steps, "loss_readback": [float(x) for x in loss_host],
            "h2d_gbs_rank0_alone": h2d_alone, "h2d_gbs_per_rank_all_ranks_copying": h2d_concurrent,
            "h2d_gbs_aggregate_all_ranks_copying": h2d_concurrent * world,
            "h2d_gbs_per_rank_in_step": wl.h2d_bytes / (e2e_ms / e2e_steps * 1e-3) / 1e9,
            "copy_bound_ms_per_step": wl.h2d_bytes / (h2d_concurrent * 1e9) * 1e3}


def run_extra_workload(name, rank, world, dev, timed, use_graph, steps=30):
    """Short device-resident run of another configuration; ranks whose shard is empty idle."""
    from codeps_b200 import _native
    spec = WORKLOADS[name]
    sets = 2 if name == "semkitti_b64" else INPUT_SETS
    wl = LossWorkload(name, rank, world, dev, input_sets=sets, pin=False)
    total = global_triplets(name, world)
    if wl.triplets == 0:
        run = lambda i: None
    else:
        run = make_runner(lambda i: wl.step(i), use_graph, wl.input_sets)
    ms, _ = timed(run, steps, 3)
    out = {"value": total * steps / (ms * 1e-3), "unit": "triplets/s", "ms_per_step": ms / steps, "steps": steps,
           "global_batch": total, "per_gpu_batch_rank0": wl.triplets,
           "scaling": "strong" if spec["mode"] == "global" else "weak", "description": spec["desc"]}
    if wl.triplets:
        _native.profile_enable(True)
        for i in range(5):
            wl.step(i)
        torch.cuda.synchronize()
        prof = _native.profile_read()
        _native.profile_enable(False)
        out["photo_kernel_ms_per_launch"] = prof["photo"][0] / max(prof["photo"][1], 1)
    del wl
    torch.cuda.empty_cache()
    return out


def run_torch_cuda_eager(workload, rank, world, dev, timed, our_value, steps=5):
    """The reference algorithm (oracle port = the reference's ATen op sequence) in torch eager on the
    GPU, same per-GPU batch: what the unmodified CoDEPS loss costs on a B200 (informational)."""
    step, total = reference_step_fn(workload, str(dev), rank, world)
    ms, _ = timed(lambda i: step(), steps, 3)
    rate = total * world * steps / (ms * 1e-3)
    return {"value": rate, "unit": "triplets/s", "ms_per_step": ms / steps, "steps": steps, "per_gpu_batch": total,
            "speedup_of_codeps_b200": our_value / rate,
            "what": "oracle/photo_oracle.py (same ATen ops as algos/depth.py + misc/image_warper.py) fwd+bwd, torch "
                    "eager on cuda, fp32, including its host-side launch overhead"}


if __name__ == "__main__":
    main()
