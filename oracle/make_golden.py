"""Generate tests/golden/*.npz by running the UNMODIFIED CoDEPS reference on seeded inputs.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python oracle/make_golden.py

The reference has no golden vectors of its own, so these fixtures are the
parity pin: outputs of the reference's own ``ReconstructionLoss`` / ``SSIMLoss`` /
``EdgeAwareSmoothnessLoss`` / ``ImageWarper`` / ``CameraModel`` classes (torch CPU, fp32, plus
an fp64 run used as tie arbiter) together with the exact inputs, so that both the oracle
(tests/test_oracle_golden.py) and the CUDA kernels (tests/test_gpu_golden.py) can be checked
against the reference without the reference being present.

Two modules that are imported by the reference package ``__init__`` files but are not on the
hot path (``yacs.config``, ``skimage.exposure``) are absent from this image and are stubbed.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
sys.path.insert(0, REPO)


def import_reference():
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not found: fixtures can only be generated in the build container")
    sys.path.insert(0, REFERENCE)
    yacs_cfg = types.ModuleType("yacs.config")
    yacs_cfg.CfgNode = type("CfgNode", (dict,), {})
    sys.modules.setdefault("yacs", types.ModuleType("yacs"))
    sys.modules.setdefault("yacs.config", yacs_cfg)
    exposure = types.ModuleType("skimage.exposure")
    exposure.match_histograms = exposure.is_low_contrast = None
    sys.modules.setdefault("skimage", types.ModuleType("skimage"))
    sys.modules.setdefault("skimage.exposure", exposure)
    import algos.depth as ref_depth  # noqa: E402
    import misc as ref_misc  # noqa: E402
    return ref_depth, ref_misc


def run_reference(ref_depth, ref_misc, batch, num_scales, seed, dtype, noise=None):
    """One forward+backward of the reference classes; returns outputs as numpy arrays.
    ``noise`` (fp64 run only) replays the fp32 run's tie-break draws: ``torch.randn`` yields
    different numbers in fp64 for the same seed, and the arbiter must see the same inputs."""
    w, h = batch.width, batch.height
    cast = lambda t: t.detach().clone().to(dtype)
    images = tuple(cast(i) for i in batch.images)
    depth = cast(batch.depth).requires_grad_(True)
    disp = cast(batch.disp).requires_grad_(True)
    poses = [cast(p).requires_grad_(True) for p in batch.poses]
    cams = [ref_misc.CameraModel.from_tensor(w, h, k) for k in batch.intrinsics]

    recon_fn = ref_depth.ReconstructionLoss(w, h, ref_depth.SSIMLoss(), num_scales,
                                            torch.device("cpu"))
    if dtype == torch.float64:
        # the reference builds its pixel grids with .float() (misc/image_warper.py:62-66);
        # promote them so the fp64 run is a true fp64 arbiter.
        for warper in recon_fn.image_warpers.values():
            i2p = warper.coordinate_warper.image_to_pointcloud
            i2p.u2d_vals = i2p.u2d_vals.double()
            i2p.v2d_vals = i2p.v2d_vals.double()

    captured = []
    real_min = torch.min

    def recording_min(*args, **kwargs):
        out = real_min(*args, **kwargs)
        if len(args) >= 1 and torch.is_tensor(args[0]) and args[0].dim() == 4 and \
                kwargs.get("dim", args[1] if len(args) > 1 else None) == 1:
            captured.append((args[0].detach().clone(), out[1].detach().clone()))
        return out

    default = torch.get_default_dtype()
    torch.set_default_dtype(dtype)  # randn / zeros at algos/depth.py:270,317 follow the default
    torch.min = recording_min
    real_randn = torch.randn
    if noise is not None:
        replay = iter(noise)
        torch.randn = lambda *a, **k: next(replay).to(dtype)
    try:
        torch.manual_seed(seed)
        recon = recon_fn(cams, images, depth, poses)
    finally:
        torch.min = real_min
        torch.randn = real_randn
        torch.set_default_dtype(default)
    smooth = ref_depth.EdgeAwareSmoothnessLoss()(images[0], disp)
    (recon + smooth).backward()
    assert len(captured) == num_scales
    out = {
        "recon": recon.detach().numpy(),
        "smooth": smooth.detach().numpy(),
        "grad_depth": depth.grad.numpy(),
        "grad_disp": disp.grad.numpy(),
        "grad_pose0": poses[0].grad.numpy(),
        "grad_pose1": poses[1].grad.numpy(),
    }
    for s, (cands, which) in enumerate(captured):
        out[f"cand{s}"] = cands.numpy()
        out[f"argmin{s}"] = which.numpy().astype(np.uint8)
    return out


def standalone_ops(ref_depth, ref_misc, batch):
    """Outputs of the individually exported operators at full resolution."""
    w, h = batch.width, batch.height
    cams = [ref_misc.CameraModel.from_tensor(w, h, k) for k in batch.intrinsics]
    warper = ref_misc.ImageWarper(w, h, torch.device("cpu"))
    depth = batch.depth.clone().requires_grad_(True)
    pose = batch.poses[1].clone().requires_grad_(True)
    grid = warper.coordinate_warper(cams, depth, pose)
    warped = warper(cams, batch.images[2], depth, pose)
    nearest = warper(cams, batch.images[2], depth.detach(), pose.detach(), interp_mode="nearest")
    upstream = torch.linspace(-1, 1, warped.numel()).view_as(warped).roll(7)
    (warped * upstream).sum().backward()
    x = batch.images[1].clone().requires_grad_(True)
    y = batch.images[0].clone().requires_grad_(True)
    ssim = ref_depth.SSIMLoss()(x, y)
    (ssim * upstream).sum().backward()
    return {
        "op_grid": grid.detach().numpy(), "op_warped": warped.detach().numpy(),
        "op_nearest": nearest.numpy(), "op_upstream": upstream.numpy(),
        "op_warp_grad_depth": depth.grad.numpy(), "op_warp_grad_pose": pose.grad.numpy(),
        "op_ssim": ssim.detach().numpy(), "op_ssim_grad_x": x.grad.numpy(),
        "op_ssim_grad_y": y.grad.numpy(),
    }


def heads_fixture(out_dir):
    """Outputs of the reference's PoseHead.transformation_from_parameters / DepthHead.disp_to_depth."""
    from models.depth_head import DepthHead
    from models.pose_head import PoseHead
    gen = torch.Generator().manual_seed(77)
    aa = (0.05 * torch.randn(6, 1, 3, generator=gen))
    aa[4] = 0.0  # zero rotation: norm backward is 0 at the origin
    tr = 0.3 * torch.randn(6, 1, 3, generator=gen)
    up = torch.randn(6, 4, 4, generator=gen)
    disp = torch.rand(2, 1, 12, 20, generator=gen)
    up_d = torch.randn(2, 1, 12, 20, generator=gen)
    blob = {"axisangle": aa.numpy(), "translation": tr.numpy(), "upstream": up.numpy(), "disp": disp.numpy(),
            "upstream_depth": up_d.numpy()}
    for invert in (False, True):
        a = aa.clone().double().requires_grad_(True)
        t = tr.clone().double().requires_grad_(True)
        m = PoseHead.transformation_from_parameters(a, t, invert)
        (m * up.double()).sum().backward()
        tag = "inv" if invert else "fwd"
        blob[f"T_{tag}"] = m.detach().numpy()
        blob[f"grad_axisangle_{tag}"] = a.grad.numpy()
        blob[f"grad_translation_{tag}"] = t.grad.numpy()
        blob[f"T32_{tag}"] = PoseHead.transformation_from_parameters(aa.clone(), tr.clone(), invert).numpy()
    d = disp.clone().double().requires_grad_(True)
    depth = DepthHead.disp_to_depth(d)
    (depth * up_d.double()).sum().backward()
    blob["depth"] = depth.detach().numpy()
    blob["grad_disp"] = d.grad.numpy()
    path = os.path.join(out_dir, "heads.npz")
    np.savez_compressed(path, **blob)
    print(f"heads -> {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


def c2c_fixture(ref_misc, out_dir):
    """Outputs of the reference's Mixup.warp_c2c (datasets/mixup.py:211-229) for every mode its
    callers use.  ``kornia`` (imported by datasets/mixup.py for an unrelated function) is absent
    from this image and stubbed."""
    kornia_contrib = types.ModuleType("kornia.contrib")
    kornia_contrib.distance_transform = None
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))
    sys.modules.setdefault("kornia.contrib", kornia_contrib)
    from datasets.mixup import Mixup
    gen = torch.Generator().manual_seed(55)
    b, hs, ws, ht, wt = 2, 24, 40, 30, 52   # source and target images differ in size
    img_src = torch.rand(b, 3, hs, ws, generator=gen)
    lbl_src = torch.randint(0, 19, (b, hs, ws), generator=gen)      # [B,H,W] int64 label map
    img_tgt = torch.rand(b, 3, ht, wt, generator=gen)
    # intrinsics as CameraModel.from_tensor delivers them (np.float32), per sample different
    k_src = torch.tensor([[45.3, 44.1, 19.7, 12.2], [38.9, 40.2, 21.4, 10.8]])
    k_tgt = torch.tensor([[41.7, 43.9, 25.1, 15.6], [60.5, 58.3, 27.9, 13.3]])
    cams_src = [ref_misc.CameraModel.from_tensor(ws, hs, k) for k in k_src]
    cams_tgt = [ref_misc.CameraModel.from_tensor(wt, ht, k) for k in k_tgt]
    blob = {"img_src": img_src.numpy(), "lbl_src": lbl_src.numpy(), "k_src": k_src.numpy(), "k_tgt": k_tgt.numpy(),
            "out_hw": np.array([ht, wt])}
    for interp in ("bilinear", "nearest"):
        for pad in ("border", "zeros"):
            blob[f"img_{interp}_{pad}"] = Mixup.warp_c2c(cams_src, cams_tgt, img_src, img_tgt, interp_mode=interp,
                                                         padding_mode=pad).numpy()
            blob[f"lbl_{interp}_{pad}"] = Mixup.warp_c2c(cams_src, cams_tgt, lbl_src, img_tgt, interp_mode=interp,
                                                         padding_mode=pad).numpy()
    blob["img_depth7"] = Mixup.warp_c2c(cams_src, cams_tgt, img_src, img_tgt, depth_val=7.5).numpy()
    path = os.path.join(out_dir, "c2c.npz")
    np.savez_compressed(path, **blob)
    print(f"c2c -> {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


def metrics_fixture(out_dir):
    """Outputs of the reference's DepthEvaluator (eval/depth.py) on seeded sparse ground truth."""
    from eval.depth import DepthEvaluator
    gen = torch.Generator().manual_seed(63)
    b, h, w = 3, 37, 61
    pred = 0.5 + 60 * torch.rand(b, 1, h, w, generator=gen)**2
    gt = pred[:, 0] * (1.3 + 0.25 * torch.randn(b, h, w, generator=gen))      # off-scale by ~1.3
    gt[torch.rand(b, h, w, generator=gen) < 0.6] = 0.0                         # sparse (LiDAR-like), VOID = 0
    gt[1, :20] = 0.0
    gt = gt.clamp(min=0.0)
    sem = torch.randint(0, 5, (b, h, w), generator=gen)
    sem[torch.rand(b, h, w, generator=gen) < 0.1] = 255
    sem[sem == 3] = 2
    gt[sem == 4] = 0.0                                                         # a class without ground truth
    blob = {"gt": gt.numpy(), "pred": pred.numpy(), "sem": sem.numpy()}
    configs = {"scaled": (True, (0.1, 80.0), False), "raw": (False, (1.0, 50.0), False),
               "garg": (True, (0.001, 80.0), True)}
    for name, (scale, rng, garg) in configs.items():
        ev = DepthEvaluator(scale, rng, garg)
        for k, v in ev.compute_depth_metrics(gt.clone(), pred.clone()).items():
            blob[f"{name}_{k}"] = v.numpy()
        if not garg:
            for k, v in ev.compute_depth_metrics_per_class(gt.clone(), pred.clone(), sem.clone()).items():
                blob[f"{name}_class_{k}"] = v.numpy()
    path = os.path.join(out_dir, "metrics.npz")
    np.savez_compressed(path, **blob)
    print(f"metrics -> {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


def flow_fixture(ref_depth, out_dir):
    """Outputs of the reference's FlowSmoothnessLoss / FlowSparsityLoss (fp64 arbiter + fp32)."""
    gen = torch.Generator().manual_seed(91)
    maps = [0.05 * torch.randn(2, 3, 13, 22, generator=gen) for _ in range(2)]
    maps[1][0, 1] = 0.0            # an all-zero plane: spatial mean 0, gradient 0
    maps[0][1, 2, 3:6, 4:9] = 0.0  # exact zeros inside a plane: sign(0) = 0
    blob = {"map0": maps[0].numpy(), "map1": maps[1].numpy()}
    cases = {"smooth_wrap": lambda m: ref_depth.FlowSmoothnessLoss(True)(m),
             "smooth_crop": lambda m: ref_depth.FlowSmoothnessLoss(False)(m),
             "sparsity": lambda m: ref_depth.FlowSparsityLoss()(m)}
    for name, fn in cases.items():
        for dtype, tag in ((torch.float64, "64"), (torch.float32, "32")):
            leaves = tuple(m.clone().to(dtype).requires_grad_(True) for m in maps)
            loss = fn(leaves)
            loss.backward()
            blob[f"{name}_loss{tag}"] = loss.detach().numpy()
            for i, leaf in enumerate(leaves):
                blob[f"{name}_grad{i}_{tag}"] = leaf.grad.numpy()
    path = os.path.join(out_dir, "flow.npz")
    np.savez_compressed(path, **blob)
    print(f"flow -> {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


CASES = {
    # name: (batch, W, H, intrinsics@WxH, num_scales, data seed, noise seed, kwargs)
    "city_near": (2, 96, 48, (106.06, 106.19, 51.42, 24.05), 5, 11, 1234, dict(depth_range="near")),
    "kitti_odd": (2, 132, 70, (51.8, 102.87, 63.94, 44.45), 5, 12, 99,
                  dict(depth_range="near", flip_every_other=True)),
    "city_wide": (1, 64, 32, (70.7, 70.8, 34.3, 16.0), 4, 13, 7, dict(depth_range="wide")),
}


def semantic_fixture(ref_depth, ref_misc, out_dir):
    """ReconstructionLoss.__call__ with semantic_mask (algos/depth.py:284-292,307-308): label maps of
    the three frames, nearest resize + nearest warp, SSIM + L1 of the label values, no auto-mask."""
    from codeps_b200.synthetic import make_batch
    b, w, h, scales = 2, 96, 64, 3
    batch = make_batch(b, w, h, (100.0, 101.0, 47.0, 31.5), seed=77, shift_px=2, flip_every_other=True)
    gen = torch.Generator().manual_seed(78)
    # piecewise-constant label maps (blocks of 8x8 pixels), shifted like the images
    coarse = torch.randint(0, 19, (b, h // 8, w // 8), generator=gen)
    lbl_t = coarse.repeat_interleave(8, 1).repeat_interleave(8, 2)
    labels = (lbl_t, torch.roll(lbl_t, 2, dims=2), torch.roll(lbl_t, -2, dims=2))
    cams = [ref_misc.CameraModel.from_tensor(w, h, k) for k in batch.intrinsics]
    loss_fn = ref_depth.ReconstructionLoss(w, h, ref_depth.SSIMLoss(), scales, torch.device("cpu"))
    loss = loss_fn(cams, batch.images, batch.depth, list(batch.poses), None, labels)
    blob = {"width": w, "height": h, "num_scales": scales, "depth": batch.depth.numpy(), "pose0": batch.poses[0].numpy(),
            "pose1": batch.poses[1].numpy(), "intrinsics": batch.intrinsics.numpy(),
            "labels": torch.stack(labels).numpy(), "loss": np.float64(loss.item())}
    path = os.path.join(out_dir, "semantic.npz")
    np.savez_compressed(path, **blob)
    print(f"semantic -> {path} loss={loss.item():.9f} ({os.path.getsize(path) / 1e3:.1f} kB)")


def main():
    from codeps_b200.synthetic import make_batch
    from oracle.photo_oracle import draw_noise
    ref_depth, ref_misc = import_reference()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order for the committed numbers
    heads_fixture(out_dir)
    flow_fixture(ref_depth, out_dir)
    c2c_fixture(ref_misc, out_dir)
    metrics_fixture(out_dir)
    semantic_fixture(ref_depth, ref_misc, out_dir)
    if "--heads-only" in sys.argv or "--small-only" in sys.argv:
        return
    for name, (b, w, h, k, scales, seed, noise_seed, kw) in CASES.items():
        batch = make_batch(b, w, h, k, seed=seed, **kw)
        blob = {
            "width": w, "height": h, "num_scales": scales, "noise_seed": noise_seed,
            "tgt": batch.images[0].numpy(), "prev": batch.images[1].numpy(),
            "next": batch.images[2].numpy(), "disp": batch.disp.numpy(),
            "depth": batch.depth.numpy(), "pose0": batch.poses[0].numpy(),
            "pose1": batch.poses[1].numpy(), "intrinsics": batch.intrinsics.numpy(),
        }
        noise = draw_noise(b, w, h, scales, noise_seed)
        for s, n in enumerate(noise):
            blob[f"noise{s}"] = n.numpy()
        ref32 = run_reference(ref_depth, ref_misc, batch, scales, noise_seed, torch.float32)
        ref64 = run_reference(ref_depth, ref_misc, batch, scales, noise_seed, torch.float64, noise)
        blob.update({f"ref32_{k_}": v for k_, v in ref32.items()})
        # fp64: keep what the tie arbiter and tolerance checks need, in fp64
        for key in ("recon", "smooth", "grad_depth", "grad_disp", "grad_pose0", "grad_pose1"):
            blob[f"ref64_{key}"] = ref64[key]
        for s in range(scales):
            top2 = np.sort(ref64[f"cand{s}"], axis=1)[:, :2]
            blob[f"ref64_gap{s}"] = (top2[:, 1] - top2[:, 0]).astype(np.float32)
            blob[f"ref64_argmin{s}"] = ref64[f"argmin{s}"]
        if name == "city_near":
            blob.update(standalone_ops(ref_depth, ref_misc, batch))
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: recon={float(ref32['recon']):.9f} smooth={float(ref32['smooth']):.9f} "
              f"hist0={np.bincount(ref32['argmin0'].ravel(), minlength=4).tolist()} "
              f"-> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
