#!/usr/bin/env python
"""bench.py -- photometric-loss fwd+bwd throughput of the CoDEPS hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
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

WORKLOADS = {
    # name: (preset, per-GPU batch, description)
    "cityscapes_b8": ("cityscapes", 8, "Cityscapes-shaped 1024x512 triplets, batch 8 per GPU, 5 scales"),
    "kitti360_b8": ("kitti360", 8, "KITTI-360-shaped 1408x376 triplets, batch 8 per GPU, per-sample intrinsics"),
    "semkitti_b8": ("semkitti", 8, "SemKITTI-DVPS-shaped 1280x384 triplets, batch 8 per GPU"),
}
RECON_WEIGHT, SMOOTH_WEIGHT = 10.0, 0.001  # cfg/train_cityscapes.yaml:40-41
NUM_SCALES = 5
INPUT_SETS = 3  # rotating input sets so that consecutive steps do not hit L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cityscapes_b8")
    ap.add_argument("--noise", choices=("torch", "fused"), default="torch")
    ap.add_argument("--intrinsics", choices=("host", "device"), default="host",
                    help="host: CameraModel objects hold host values (kernel parameter space); device: lazy "
                         "CameraModel.from_tensor of CUDA rows, the kernels read the calibration from HBM")
    ap.add_argument("--overlap-smooth", action="store_true",
                    help="run the smoothness loss on a second stream next to the reconstruction loss")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def profiled_kernel(workload: str, kernel_substr: str):
    """Latest tracked ncu record (profiles/*_ncu.json, written by tools/ncu_to_profile.py) of a
    kernel for this workload: DRAM traffic per launch, issue-slot utilisation, L2 hit rate."""
    import glob
    best = None
    import re

    def natural(path):  # r01_9_... before r01_11_...
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", os.path.basename(path))]
    for path in sorted(glob.glob(os.path.join(REPO, "profiles", "*_ncu.json")), key=natural):
        try:
            with open(path) as f:
                blob = json.load(f)
        except (OSError, ValueError):
            continue
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


# ------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle port of the reference torch path on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(preset: str, steps: int, warmup: int, budget_s: float = 25.0):
    """fwd+bwd of the reference algorithm (oracle/photo_oracle.py, same ATen ops as the
    reference) on ONE triplet of the workload per step, all host threads.  Returns
    (triplets/s, cores, iterations, median seconds)."""
    from codeps_b200 import synthetic
    from oracle import photo_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tb = synthetic.make_preset_batch(preset, 1, seed=1000)
    noise = po.draw_noise(1, tb.width, tb.height, NUM_SCALES, seed=1)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        po.loss_and_grads(tb.intrinsics.numpy(), tb.images, tb.depth, tb.disp, tb.poses, noise, NUM_SCALES,
                          recon_weight=RECON_WEIGHT, smooth_weight=SMOOTH_WEIGHT)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if i >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    med = statistics.median(times)
    return 1.0 / med, cores, len(times), med


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from codeps_b200 import synthetic
    preset, batch, desc = WORKLOADS[args.workload]
    w, h = synthetic.PRESETS[preset][0], synthetic.PRESETS[preset][1]
    rate, cores, iters, med = cpu_reference_rate(preset, args.steps, max(args.warmup, 1), budget_s=120.0)
    sample = f"1 of {batch} triplets of the batch per step ({iters} steps timed), oracle port of the reference torch-CPU path"
    line = {
        "impl": "reference", "metric": f"photometric-loss fwd+bwd frame-triplets/sec @{w}x{h}", "value": rate,
        "unit": "triplets/s", "n_gpus": args.gpus, "steps": iters, "warmup": max(args.warmup, 1),
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "per_gpu_batch": batch, "num_scales": NUM_SCALES,
                   "device": "host CPU", "torch_threads": cores},
        "cpu_baseline": {"value": rate, "unit": "triplets/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
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

    preset, batch, desc = WORKLOADS[args.workload]
    w, h = synthetic.PRESETS[preset][0], synthetic.PRESETS[preset][1]
    flip = preset != "cityscapes"

    # ---- inputs: INPUT_SETS distinct batches per rank, pinned on the host and resident in HBM
    host_sets = [synthetic.make_preset_batch(preset, batch, seed=1000 * rank + i, flip_every_other=flip).pin()
                 for i in range(INPUT_SETS)]
    dev_sets = [hs.to(dev) for hs in host_sets]
    cams = [hs.camera_models() for hs in host_sets]
    if args.intrinsics == "device":
        k_dev = [hs.intrinsics.to(dev) for hs in host_sets]
        cams = [[codeps_b200.CameraModel.from_tensor(w, h, k[i]) for i in range(batch)] for k in k_dev]
    resident_bytes = sum(hs.nbytes() for hs in host_sets)

    recon_fn = codeps_b200.ReconstructionLoss(w, h, codeps_b200.SSIMLoss(), NUM_SCALES, dev, noise=args.noise)
    smooth_fn = codeps_b200.EdgeAwareSmoothnessLoss()
    w_recon = torch.tensor(RECON_WEIGHT, device=dev)
    w_smooth = torch.tensor(SMOOTH_WEIGHT, device=dev)

    side_stream = torch.cuda.Stream() if args.overlap_smooth else None

    def make_step(loss_fn):
        def step(i):
            """fwd + bwd on resident input set i; returns (recon, smooth, grads)."""
            ds = dev_sets[i]
            # fresh autograd leaves every step (views, no copies), created on the launching stream
            depth, disp = ds.depth.detach().requires_grad_(True), ds.disp.detach().requires_grad_(True)
            p0, p1 = ds.poses[0].detach().requires_grad_(True), ds.poses[1].detach().requires_grad_(True)
            if side_stream is not None:
                # the two losses are independent: the smoothness kernels run on a second stream and
                # fill the SMs the tile kernel's last wave and the small reduction kernels leave idle
                cur = torch.cuda.current_stream()
                side_stream.wait_stream(cur)
                with torch.cuda.stream(side_stream):
                    smooth = smooth_fn(ds.images[0], disp)
                recon = loss_fn(cams[i], ds.images, depth, (p0, p1))
                cur.wait_stream(side_stream)
            else:
                recon = loss_fn(cams[i], ds.images, depth, (p0, p1))
                smooth = smooth_fn(ds.images[0], disp)
            # the caller's  loss = 10*recon + 0.001*smooth; loss.backward()  (train_codeps.py:102-107)
            grads = torch.autograd.grad([recon, smooth], [depth, disp, p0, p1], grad_outputs=[w_recon, w_smooth])
            return recon, smooth, grads
        return step

    def make_runner(step_fn, use_graph):
        """Warm up, optionally capture one CUDA graph per input set, return run(i)."""
        for i in range(3):
            step_fn(i % INPUT_SETS)
        torch.cuda.synchronize()
        if not use_graph:
            return lambda i: step_fn(i % INPUT_SETS)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(INPUT_SETS):
                step_fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs, outs = [], []
        for i in range(INPUT_SETS):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs.append(step_fn(i))
            graphs.append(g)

        def run(i):
            graphs[i % INPUT_SETS].replay()
            return outs[i % INPUT_SETS]
        return run

    step = make_step(recon_fn)
    use_graph = not args.no_graph
    torch.manual_seed(1234 + rank)
    for i in range(3):  # eager warm-up (sizes the allocator pools, sets kernel attributes)
        out = step(i % INPUT_SETS)
    torch.cuda.synchronize()
    launches_before = ops.launch_count()
    out = step(0)
    launches_per_step = ops.launch_count() - launches_before
    torch.cuda.synchronize()
    run_step = make_runner(step, use_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run, steps, sampler=None):
        """W warm-up steps, then exactly `steps` steps between barriers; device ms, max over ranks."""
        for i in range(max(args.warmup, 3)):
            run(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        last = None
        for i in range(steps):
            last = run(i)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), last

    # ---- timed region: device-resident inputs
    with ClockSampler(local_rank) as clocks:
        elapsed_ms, last = timed(run_step, args.steps)
    value = n_gpus * batch * args.steps / (elapsed_ms * 1e-3)
    recon_val, smooth_val = float(last[0].detach()), float(last[1].detach())

    # ---- eager pass with per-kernel CUDA events (same K steps) -> dominant-kernel duration
    _native.profile_enable(True)
    for i in range(args.steps):
        step(i % INPUT_SETS)
    torch.cuda.synchronize()
    prof = _native.profile_read()
    _native.profile_enable(False)
    kernel_ms = {k: (ms / n if n else None) for k, (ms, n) in prof.items()}
    photo_ms = kernel_ms["photo"]
    s0 = sum(ww * hh for ww, hh in synthetic.level_sizes(w, h, NUM_SCALES))
    a_alg = synthetic.algorithmic_bytes(w, h, NUM_SCALES)
    # fused tile kernel = the per-level forward (41 B/level-px) and backward (45 B/level-px) of
    # SURVEY.md section 8d done in one pass
    photo_alg_bytes = 86 * s0 * batch
    peak, peak_src = measured_peaks()
    achieved = photo_alg_bytes / (photo_ms * 1e-3) / 1e9
    step_kernel_ms = sum(v for v in kernel_ms.values() if v)
    prof_rec = profiled_kernel(args.workload, "cdp_photo_kernel<1") or profiled_kernel(args.workload, "cdp_photo_kernel")
    roofline = {
        "bound": "hbm", "kernel": "cdp_photo_kernel<true,false>", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": prof_rec["dram_bytes"] if prof_rec else None, "peak_source": peak_src,
        "ncu": ({k: prof_rec.get(k) for k in ("profile", "duration_us", "issue_slot_pct", "sm_pct_of_peak",
                                              "dram_pct_of_peak", "l2_hit_pct", "l1_hit_pct",
                                              "achieved_occupancy_pct", "registers_per_thread")}
                if prof_rec else None),
        "limiter": "fp32 issue slots, not HBM (DESIGN.md section 6): see ncu.issue_slot_pct vs ncu.dram_pct_of_peak",
        "algorithmic_bytes_per_launch": photo_alg_bytes, "kernel_ms": photo_ms,
        "kernel_share_of_step": photo_ms / step_kernel_ms if step_kernel_ms else None,
        "kernel_ms_all": kernel_ms,
        "path_bytes_per_triplet": a_alg,
        "path_frac": (value / n_gpus) * a_alg / 1e9 / peak,
    }

    # ---- same steps with the kernel's counter-based tie-break noise instead of torch.randn per level
    extras = {}
    if args.noise == "torch":
        fused_fn = codeps_b200.ReconstructionLoss(w, h, codeps_b200.SSIMLoss(), NUM_SCALES, dev, noise="fused")
        fused_ms, _ = timed(make_runner(make_step(fused_fn), use_graph), args.steps)
        extras["value_fused_noise"] = n_gpus * batch * args.steps / (fused_ms * 1e-3)
        extras["note"] = ("value_fused_noise: same step with ReconstructionLoss(noise='fused') -- no torch.randn "
                          "launches / noise traffic; different random numbers than the reference's stream")

    # ---- end to end: host (pinned) inputs -> public classes -> loss read back on the host
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        main_stream = torch.cuda.current_stream()
        e2e_steps = max(10, min(args.steps, 40))

        def upload(i):
            with torch.cuda.stream(copy_stream):
                dsb = host_sets[i % INPUT_SETS].to(dev, non_blocking=True)
                evt = torch.cuda.Event()
                evt.record(copy_stream)
            return dsb, evt

        loss_host = torch.zeros(2, pin_memory=True)

        def e2e_step(i, staged):
            dsb, evt = staged
            nxt = upload(i + 1)  # overlap the next step's H2D with this step's kernels
            main_stream.wait_event(evt)
            depth = dsb.depth.requires_grad_(True)
            disp = dsb.disp.requires_grad_(True)
            p0, p1 = dsb.poses[0].requires_grad_(True), dsb.poses[1].requires_grad_(True)
            recon = recon_fn(cams[i % INPUT_SETS], dsb.images, depth, (p0, p1))
            smooth = smooth_fn(dsb.images[0], disp)
            loss = RECON_WEIGHT * recon + SMOOTH_WEIGHT * smooth
            loss.backward()
            loss_host.copy_(torch.stack((recon.detach(), smooth.detach())), non_blocking=True)
            for tns in list(dsb.images) + [dsb.depth, dsb.disp] + list(dsb.poses):
                tns.record_stream(main_stream)
            return nxt

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
        e2e_ms = e0.elapsed_time(e1)
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": n_gpus * batch * e2e_steps / (e2e_ms * 1e-3), "unit": "triplets/s",
               "h2d_bytes_per_step": host_sets[0].nbytes(), "d2h_bytes_per_step": 8, "steps": e2e_steps,
               "ms_per_step": e2e_ms / e2e_steps, "loss_readback": [float(x) for x in loss_host]}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        rate, cores, iters, med = cpu_reference_rate(preset, steps=12, warmup=1, budget_s=20.0)
        cpu_baseline = {"value": rate, "unit": "triplets/s", "cores": cores, "kind": "port",
                        "sample": f"1 triplet of the workload per iteration (B=1, {w}x{h}), {iters} iterations, "
                                  f"median {med * 1e3:.0f} ms, torch {torch.__version__} CPU"}

    # scalar statistics only (one coalesced all-reduce of 3 doubles); the data path has no collective
    from codeps_b200.distributed import reduce_loss_dict
    stats = reduce_loss_dict({"recon": torch.tensor(recon_val, device=dev, dtype=torch.float64),
                              "smooth": torch.tensor(smooth_val, device=dev, dtype=torch.float64)}, batch)
    checksum = [float(stats["recon"]), float(stats["smooth"])]

    if rank == 0:
        line = {
            "metric": f"photometric-loss fwd+bwd frame-triplets/sec @{w}x{h}",
            "value": value, "unit": "triplets/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "per_gpu_batch": batch,
                       "global_batch": batch * n_gpus, "num_scales": NUM_SCALES, "noise": args.noise, "intrinsics": args.intrinsics, "overlap_smooth": bool(args.overlap_smooth),
                       "timed_with": "cuda_graph_replay" if use_graph else "eager_launches",
                       "l2": f"{INPUT_SETS} rotating input sets, {resident_bytes / 1e6:.0f} MB resident > 126 MB L2",
                       "loss_weights": [RECON_WEIGHT, SMOOTH_WEIGHT]},
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


if __name__ == "__main__":
    main()
