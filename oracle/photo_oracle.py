"""CPU oracle for the photometric-loss hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, op for op, what the CoDEPS reference computes on the path
``ReconstructionLoss`` + ``SSIMLoss`` + ``EdgeAwareSmoothnessLoss`` on top of
``ImageWarper`` / ``CameraModel``.  It exists to check the CUDA kernels; it is imported
only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs, never
by the ``codeps_b200`` package (which has no CPU path at all).

Parity pin: the reference has no tests or golden vectors for this path,
but it is importable in the build container.  ``oracle/make_golden.py`` runs the *reference
itself* (imported from /root/reference) on seeded inputs and commits its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` requires this oracle to reproduce them
(bit-for-bit on the argmin, to float rounding on values).  So: parity pinned against
outputs of the reference run in the build container.  The restatements of the neighbouring
rows are pinned the same way: ``transformation_from_parameters`` /
``disp_to_depth`` by ``heads.npz`` (tests/test_heads.py), ``flow_smoothness_loss`` /
``flow_sparsity_loss`` by ``flow.npz`` (tests/test_flow_losses.py), ``warp_c2c`` by ``c2c.npz``
(tests/test_warp_c2c.py), ``depth_metrics[_per_class]`` by ``metrics.npz``
(tests/test_depth_metrics.py) -- each generated from the reference's own class.

The arithmetic all lives in PyTorch ATen ops (reference pins torch==1.12.1,
/root/reference/requirements.txt:1; validated here with torch 2.11): ``upsample_bilinear2d``
(align_corners=False), ``bmm``, ``grid_sampler_2d`` (bilinear, border, align_corners=True),
``reflection_pad2d``, ``avg_pool2d``, ``clamp``, ``min``.  The oracle calls the same ATen ops
in the same order so that in fp32 it tracks the reference to rounding, and it is
dtype-generic: run it in fp64 (``dtype=torch.float64``) to arbitrate near-ties.

Functions take intrinsics as a ``[B,4]`` array-like of (fx, fy, cx, cy) already scaled to the
image size of the call.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SSIM_C1 = 0.01**2  # /root/reference/algos/depth.py:125
SSIM_C2 = 0.03**2  # /root/reference/algos/depth.py:126
Z_MIN = 1e-5  # /root/reference/misc/image_warper.py:32


def _k_columns(intrinsics, like: torch.Tensor):
    """[B,4] intrinsics -> four [B,1,1] tensors in ``like``'s dtype (each value is first
    rounded to fp32, which is what the reference's python-scalar-meets-fp32-tensor rule does)."""
    k = torch.as_tensor(intrinsics, dtype=torch.float32).to(like.dtype).to(like.device)
    return [k[:, i].view(-1, 1, 1) for i in range(4)]


def unit_rays(intrinsics, height: int, width: int, like: torch.Tensor):
    """Unit viewing ray per pixel: /root/reference/misc/camera_model.py:52-71 evaluated on
    the pixel grid of /root/reference/misc/image_warper.py:62-66.  Returns three [B,H,W]."""
    fx, fy, cx, cy = _k_columns(intrinsics, like)
    u = torch.arange(width, dtype=like.dtype, device=like.device).view(1, 1, width)
    v = torch.arange(height, dtype=like.dtype, device=like.device).view(1, height, 1)
    rx = ((u - cx) / fx).expand(-1, height, -1)
    ry = ((v - cy) / fy).expand(-1, -1, width)
    norm = torch.sqrt(rx**2 + ry**2 + 1.0)
    return rx / norm, ry / norm, 1.0 / norm


def backproject(depth: torch.Tensor, intrinsics) -> torch.Tensor:
    """depth [B,1,H,W] -> point cloud [B,3,H,W] (image_warper.py:68-87): the ray is scaled so
    that its z component equals the depth value."""
    _, _, h, w = depth.shape
    rx, ry, rz = (r.unsqueeze(1) for r in unit_rays(intrinsics, h, w, depth))
    return torch.cat((depth / rz.abs() * rx, depth / rz.abs() * ry, depth / rz.abs() * rz), dim=1)


def reproject_grid(depth: torch.Tensor, pose: torch.Tensor, intrinsics,
                   motion: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Normalised sampling grid [B,H,W,2] (image_warper.py:100-144 and :20-51)."""
    b, _, h, w = depth.shape
    cloud = backproject(depth, intrinsics)
    homo = torch.cat((cloud, torch.ones_like(depth)), dim=1).view(b, 4, -1)
    moved = torch.bmm(pose, homo).view(b, 4, h, w)
    if motion is not None:
        moved = torch.cat((moved[:, :3] + motion, moved[:, 3:]), dim=1)
    eucl = moved[:, :3] / moved[:, 3:4]
    fx, fy, cx, cy = _k_columns(intrinsics, depth)
    z = eucl[:, 2].clamp(min=Z_MIN)
    u = (eucl[:, 0] / z) * fx + cx
    v = (eucl[:, 1] / z) * fy + cy
    gx = (u / (w - 1) - 0.5) * 2
    gy = (v / (h - 1) - 0.5) * 2
    return torch.stack((gx, gy), dim=3)


def warp_image(src: torch.Tensor, depth: torch.Tensor, pose: torch.Tensor, intrinsics,
               mode: str = "bilinear", motion: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ImageWarper.forward (image_warper.py:153-184)."""
    grid = reproject_grid(depth, pose, intrinsics, motion)
    return F.grid_sample(src, grid, mode=mode, padding_mode="border", align_corners=True)


def ssim_loss_map(x: torch.Tensor, y: torch.Tensor, return_raw: bool = False):
    """SSIMLoss.__call__ (algos/depth.py:128-155): 3x3 box statistics on reflect-padded
    inputs; returns clamp((1-SSIM)/2, 0, 1) per channel (and, on request, the un-clamped
    value, whose distance to 0 / 1 tells how close the clamp is to switching)."""
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    yp = F.pad(y, (1, 1, 1, 1), mode="reflect")
    box = lambda t: F.avg_pool2d(t, 3, 1)
    mu_x, mu_y = box(xp), box(yp)
    var_x = box(xp**2) - mu_x**2
    var_y = box(yp**2) - mu_y**2
    cov = box(xp * yp) - mu_x * mu_y
    num = (2 * mu_x * mu_y + SSIM_C1) * (2 * cov + SSIM_C2)
    den = (mu_x**2 + mu_y**2 + SSIM_C1) * (var_x + var_y + SSIM_C2)
    raw = (1 - num / den) / 2
    out = torch.clamp(raw, 0, 1)
    return (out, raw) if return_raw else out


def photometric_error(pred: torch.Tensor, tgt: torch.Tensor, alpha: float = 0.85) -> torch.Tensor:
    """ReconstructionLoss._compute_loss (algos/depth.py:221-237) -> [B,1,H,W]."""
    l1 = (pred - tgt).abs().mean(1, True)
    ssim = ssim_loss_map(pred, tgt).mean(1, True)
    return alpha * ssim + (1 - alpha) * l1


def resize_bilinear(x: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """Interpolate(..., "bilinear") (algos/depth.py:158-173): always from full resolution."""
    return F.interpolate(x, (height, width), mode="bilinear", align_corners=False)


def scaled_intrinsics(intrinsics, full_wh: Tuple[int, int], level_wh: Tuple[int, int]):
    """camera_model.py:36-41 applied to a [B,4] fp32 table: focal length and principal point
    scale with the ratio of image sizes (python-float ratio, fp32 product)."""
    import numpy as np
    k = np.asarray(intrinsics, dtype=np.float32)
    su = level_wh[0] / full_wh[0]
    sv = level_wh[1] / full_wh[1]
    out = np.empty_like(k)
    out[:, 0] = k[:, 0] * np.float32(su)
    out[:, 1] = k[:, 1] * np.float32(sv)
    out[:, 2] = k[:, 2] * np.float32(su)
    out[:, 3] = k[:, 3] * np.float32(sv)
    return out


def reconstruction_loss(intrinsics, images: Sequence[torch.Tensor], depth: torch.Tensor,
                        poses: Sequence[torch.Tensor], noise: Sequence[torch.Tensor],
                        num_scales: int = 5, alpha: float = 0.85,
                        motions: Optional[Sequence[torch.Tensor]] = None,
                        level_intrinsics: Optional[Sequence] = None, details: bool = False,
                        forced_argmin: Optional[Sequence[torch.Tensor]] = None):
    """ReconstructionLoss.__call__ (algos/depth.py:239-326), auto-mask branch.

    ``noise[s]`` is the [B,2,H_s,W_s] tensor the reference would draw with ``torch.randn`` at
    level s (depth.py:317) -- *unscaled*; it is multiplied by 1e-5 here.  ``level_intrinsics``
    optionally overrides the per-level [B,4] tables (used when the caller has CameraModel
    objects and wants the exact python-side rounding).  With ``details`` also returns, per
    level, the stacked candidate losses [B,4,H_s,W_s], the argmin [B,H_s,W_s] (index 0/1 =
    reprojection from t-1/t+1, 2/3 = identity, i.e. auto-masked), which the reference
    computes and discards (depth.py:323), the two normalised sampling grids, and the smallest
    |warped - target| over sources and channels (distance of the L1 term to its sign switch).

    ``forced_argmin`` (test aid): per level a [B,H_s,W_s] index map to select instead of the
    minimum.  Gradients are a smooth function of the inputs only for a *given* selection; at a
    near-tie, fp32 and fp64 evaluations of the reference itself pick different candidates and
    their gradients differ at those pixels.  Forcing the selection of the implementation under
    test lets a gradient comparison ignore exactly that ambiguity (the selection itself is
    checked separately, bit-exact away from ties).
    """
    _, _, h, w = depth.shape
    total = torch.zeros(1, dtype=depth.dtype, device=depth.device)
    per_level = []
    for s in range(num_scales):
        ws, hs = w // 2**s, h // 2**s
        k_s = (level_intrinsics[s] if level_intrinsics is not None else scaled_intrinsics(
            intrinsics, (w, h), (ws, hs)))
        tgt_s = resize_bilinear(images[0], hs, ws)
        depth_s = resize_bilinear(depth, hs, ws)
        cands, grids, l1_margin, clamp_margin = [], [], None, None
        for i, frame in enumerate(images[1:]):
            src_s = resize_bilinear(frame, hs, ws)
            motion_s = None if motions is None else resize_bilinear(motions[i], hs, ws)
            grid = reproject_grid(depth_s, poses[i], k_s, motion_s)
            grids.append(grid.detach())
            warped = F.grid_sample(src_s, grid, mode="bilinear", padding_mode="border", align_corners=True)
            cands.append(photometric_error(warped, tgt_s, alpha))
            margin = (warped.detach() - tgt_s).abs().amin(1)
            l1_margin = margin if l1_margin is None else torch.minimum(l1_margin, margin)
            raw = ssim_loss_map(warped.detach(), tgt_s, return_raw=True)[1]
            margin = torch.minimum(raw.abs(), (1 - raw).abs()).amin(1)
            clamp_margin = margin if clamp_margin is None else torch.minimum(clamp_margin, margin)
        ident = [photometric_error(resize_bilinear(frame, hs, ws), tgt_s, alpha)
                 for frame in images[1:]]
        ident = torch.cat(ident, 1) + noise[s].to(depth.dtype) * 0.00001
        stacked = torch.cat(cands + [ident], dim=1)
        best, which = torch.min(stacked, dim=1)
        if forced_argmin is not None:
            which = forced_argmin[s].to(torch.int64)
            best = stacked.gather(1, which.unsqueeze(1)).squeeze(1)
        total = total + best.mean() / (2**s)
        per_level.append((stacked, which, grids, l1_margin, clamp_margin))
    loss = total[0] / num_scales
    return (loss, per_level) if details else loss


def semantic_reconstruction_loss(intrinsics, labels: Sequence[torch.Tensor], depth: torch.Tensor,
                                 poses: Sequence[torch.Tensor], num_scales: int = 5, alpha: float = 0.85,
                                 level_intrinsics=None) -> torch.Tensor:
    """The semantic_mask branch of ReconstructionLoss.__call__ (algos/depth.py:284-292,307-308,325-326):
    label maps [B,H,W] of frames t, t-1, t+1; per level nearest resize, nearest-neighbour warp of the
    two neighbouring maps, _compute_loss against the target map, mean over both and all pixels."""
    b, _, h, w = depth.shape
    total = torch.zeros(1, dtype=depth.dtype, device=depth.device)
    for s in range(num_scales):
        hs, ws = h // 2**s, w // 2**s
        k = level_intrinsics[s] if level_intrinsics is not None else scaled_intrinsics(intrinsics, (w, h), (ws, hs))
        depth_s = F.interpolate(depth, (hs, ws), mode="bilinear", align_corners=False)
        tgt = F.interpolate(labels[0].unsqueeze(1).to(depth.dtype), (hs, ws), mode="nearest")
        terms = []
        for i in range(2):
            frame = F.interpolate(labels[1 + i].unsqueeze(1).to(depth.dtype), (hs, ws), mode="nearest")
            pred = warp_image(frame, depth_s, poses[i], k, mode="nearest")
            terms.append(alpha * ssim_loss_map(pred, tgt).mean(1, True) + (1 - alpha) * (pred - tgt).abs().mean(1, True))
        total = total + torch.cat(terms, 1).mean() / 2**s
    return total[0] / num_scales


def smoothness_loss(target_image: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """EdgeAwareSmoothnessLoss.__call__ (algos/depth.py:58-107)."""
    mean_disp = disp.mean(2, True).mean(3, True)
    d = disp / (mean_disp + 1e-7)
    ddx = (d[:, :, :, :-1] - d[:, :, :, 1:]).abs()
    ddy = (d[:, :, :-1, :] - d[:, :, 1:, :]).abs()
    idx = (target_image[:, :, :, :-1] - target_image[:, :, :, 1:]).abs().mean(1, True)
    idy = (target_image[:, :, :-1, :] - target_image[:, :, 1:, :]).abs().mean(1, True)
    return (ddx * torch.exp(-idx)).mean() + (ddy * torch.exp(-idy)).mean()


def draw_noise(batch: int, width: int, height: int, num_scales: int, seed: int,
               device="cpu") -> List[torch.Tensor]:
    """The tie-break noise sequence the reference consumes (depth.py:317): one
    ``torch.randn(B,2,H_s,W_s)`` per level, in level order, from ``torch.manual_seed(seed)``."""
    torch.manual_seed(seed)
    return [torch.randn(batch, 2, height // 2**s, width // 2**s, device=device)
            for s in range(num_scales)]


def loss_and_grads(intrinsics, images, depth, disp, poses, noise, num_scales=5, alpha=0.85,
                   dtype=torch.float32, recon_weight: float = 1.0, smooth_weight: float = 1.0,
                   level_intrinsics=None, forced_argmin=None, motions=None):
    """Forward + autograd backward of (recon, smooth) on the CPU.  Returns a dict with the two
    loss values, per-level argmin/candidates and dL/d depth, dL/d disp, dL/dT_0, dL/dT_1 of
    ``recon_weight*recon + smooth_weight*smooth`` (depth and disp are treated as independent
    leaves, as at the boundary of the CUDA op)."""
    cast = lambda t: t.detach().to(dtype)
    images = [cast(i) for i in images]
    depth = cast(depth).requires_grad_(True)
    disp = cast(disp).requires_grad_(True)
    poses = [cast(p).requires_grad_(True) for p in poses]
    if motions is not None:
        motions = [cast(m).requires_grad_(True) for m in motions]
    recon, levels = reconstruction_loss(intrinsics, images, depth, poses, noise, num_scales,
                                        alpha, motions=motions, level_intrinsics=level_intrinsics, details=True,
                                        forced_argmin=forced_argmin)
    smooth = smoothness_loss(images[0], disp)
    (recon_weight * recon + smooth_weight * smooth).backward()
    return {
        "recon": recon.detach(),
        "smooth": smooth.detach(),
        "argmin": [lv[1].to(torch.uint8) for lv in levels],
        "candidates": [lv[0].detach() for lv in levels],
        "grids": [lv[2] for lv in levels],  # per level: normalised sampling grids of both sources
        "l1_margin": [lv[3] for lv in levels],  # per level: min over sources/channels |warped - target|
        "clamp_margin": [lv[4] for lv in levels],  # per level: distance of (1-SSIM)/2 to the clamp bounds
        "grad_depth": depth.grad,
        "grad_disp": disp.grad,
        "grad_pose": [p.grad for p in poses],
        "grad_motion": [m.grad for m in motions] if motions is not None else None,
    }


def rot_from_axisangle(axisangle: torch.Tensor) -> torch.Tensor:
    """PoseHead.rot_from_axisangle (/root/reference/models/pose_head.py:80-119): [B,1,3] -> [B,4,4]."""
    angle = torch.norm(axisangle, 2, 2, True)
    axis = axisangle / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    c = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xc, yc, zc = x * c, y * c, z * c
    xyc, yzc, zxc = x * yc, y * zc, z * xc
    rows = [[x * xc + ca, xyc - zs, zxc + ys], [xyc + zs, y * yc + ca, yzc - xs], [zxc - ys, yzc + xs, z * zc + ca]]
    rot = torch.zeros(axisangle.shape[0], 4, 4, dtype=axisangle.dtype, device=axisangle.device)
    for r in range(3):
        for col in range(3):
            rot[:, r, col] = rows[r][col].reshape(-1)
    rot[:, 3, 3] = 1
    return rot


def transformation_from_parameters(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool = False):
    """PoseHead.transformation_from_parameters (/root/reference/models/pose_head.py:56-77)."""
    rot = rot_from_axisangle(axisangle)
    t = -translation if invert else translation
    trans = torch.zeros(t.shape[0], 4, 4, dtype=t.dtype, device=t.device)
    for i in range(4):
        trans[:, i, i] = 1
    trans[:, :3, 3] = t.reshape(-1, 3)
    return torch.matmul(rot.transpose(1, 2), trans) if invert else torch.matmul(trans, rot)


def disp_to_depth(disp: torch.Tensor, min_depth: float = 0.1, max_depth: float = 100):
    """DepthHead.disp_to_depth (/root/reference/models/depth_head.py:49-54)."""
    min_disp, max_disp = 1 / max_depth, 1 / min_depth
    return 1 / (min_disp + (max_disp - min_disp) * disp)


def flow_smoothness_loss(flow_maps: Sequence[torch.Tensor], wrap_around: bool = True) -> torch.Tensor:
    """FlowSmoothnessLoss.__call__ (/root/reference/algos/depth.py:15-34): per map the mean of
    sqrt(dx^2 + dy^2 + 1e-7) over backward differences with wrap-around (cropped to [1:, 1:]
    otherwise), averaged over the maps."""
    total = 0
    for f in flow_maps:
        gx = f - torch.roll(f, shifts=1, dims=3)
        gy = f - torch.roll(f, shifts=1, dims=2)
        if not wrap_around:
            gx, gy = gx[:, :, 1:, 1:], gy[:, :, 1:, 1:]
        total = total + torch.sqrt(gx * gx + gy * gy + 1e-7).mean()
    return total / len(flow_maps)


def flow_sparsity_loss(flow_maps: Sequence[torch.Tensor]) -> torch.Tensor:
    """FlowSparsityLoss.__call__ (/root/reference/algos/depth.py:37-52): per map
    mean(2 m sqrt(|f| / (m + 1e-7) + 1)) with m = mean_{H,W} |f| detached, averaged over the maps."""
    total = 0
    for f in flow_maps:
        a = f.abs()
        m = a.mean(dim=(2, 3), keepdim=True).detach()
        total = total + (2 * m * torch.sqrt(a / (m + 1e-7) + 1)).mean()
    return total / len(flow_maps)


def warp_c2c(k_src, k_tgt, in_src: torch.Tensor, out_hw: Tuple[int, int], depth_val: float = 1.0,
             interp_mode: str = "bilinear", padding_mode: str = "border", ieee_sqrt: bool = False) -> torch.Tensor:
    """Mixup.warp_c2c (/root/reference/datasets/mixup.py:211-229): rays of the target camera on
    its fp32 pixel grid (misc/image_warper.py:62-87, misc/camera_model.py:52-71), point at
    ``depth_val`` in fp64, projection with the source camera (mixup.py:29-66), fp64 grid_sample.
    k_src / k_tgt: [B,4] (fx, fy, cx, cy) arrays.

    ``ieee_sqrt``: torch's vectorised CPU ``sqrt`` is not correctly rounded for fp32 (about 0.4 % of
    arguments come out 1 ulp low), CUDA's ``sqrtf`` and numpy's are.  The default reproduces the
    reference as it runs on the CPU (and matches the committed fixture to 1e-14); ``True`` takes
    the correctly rounded root, which is what the reference computes on a GPU and what the kernel
    computes."""
    if in_src.dim() == 3:
        in_src = in_src.unsqueeze(1)
    b, _, hs, ws = in_src.shape
    ht, wt = out_hw
    u = torch.arange(wt).expand(ht, wt).float()
    v = torch.arange(ht).expand(wt, ht).t().float()
    grids = []
    for i in range(b):
        fx, fy, cx, cy = (np.float32(x) for x in np.asarray(k_tgt)[i])
        rx, ry = (u - cx) / fx, (v - cy) / fy
        norm2 = rx**2 + ry**2 + 1.0
        norm = torch.from_numpy(np.sqrt(norm2.numpy())) if ieee_sqrt else torch.sqrt(norm2)
        rx, ry, rz = rx / norm, ry / norm, 1.0 / norm
        scale = torch.full((ht, wt), float(depth_val), dtype=torch.float64) / rz.double().abs()
        x3, y3, z3 = scale * rx.double(), scale * ry.double(), (scale * rz.double()).clamp(min=1e-5)
        sfx, sfy, scx, scy = (float(x) for x in np.asarray(k_src)[i])
        us, vs = (x3 / z3) * sfx + scx, (y3 / z3) * sfy + scy
        grids.append(torch.stack([(us / (ws - 1) - 0.5) * 2, (vs / (hs - 1) - 0.5) * 2], dim=-1))
    return F.grid_sample(in_src.double(), torch.stack(grids), mode=interp_mode, padding_mode=padding_mode,
                         align_corners=True)


DEPTH_STAT_KEYS = ("d_a1", "d_a2", "d_a3", "d_rmse", "d_rmse_log", "d_abs_rel", "d_sq_rel")


def _depth_stats(gt: torch.Tensor, pred: torch.Tensor):
    """DepthEvaluator._compute_depth_stats (/root/reference/eval/depth.py:108-133)."""
    thresh = torch.max(gt / pred, pred / gt)
    return {"d_a1": (thresh < 1.25).float().mean(), "d_a2": (thresh < 1.25**2).float().mean(),
            "d_a3": (thresh < 1.25**3).float().mean(), "d_rmse": torch.sqrt(((gt - pred)**2).mean()),
            "d_rmse_log": torch.sqrt(((torch.log(gt) - torch.log(pred))**2).mean()),
            "d_abs_rel": torch.mean(torch.abs(gt - pred) / gt), "d_sq_rel": torch.mean((gt - pred)**2 / gt)}


def depth_metrics(depth_gt: torch.Tensor, depth_pred: torch.Tensor, depth_range: Tuple[float, float],
                  use_gt_scale: bool, garg_crop: bool = False):
    """DepthEvaluator.compute_depth_metrics (/root/reference/eval/depth.py:21-70): per image over
    gt > 0 (and the Garg crop), median scaling, clamp, statistics; mean over the batch."""
    if depth_gt.dim() == 3:
        depth_gt = depth_gt.unsqueeze(1)
    mask = depth_gt > 0
    if garg_crop:
        h, w = depth_gt.shape[2:]
        crop = torch.zeros_like(mask)
        crop[:, :, int(0.4080 * h):int(0.9891 * h), int(0.0354 * w):int(0.9638 * w)] = 1
        mask = mask & crop
    total = {}
    for b in range(depth_gt.shape[0]):
        gt, pred = depth_gt[b][mask[b]], depth_pred[b][mask[b]]
        if use_gt_scale:
            pred = pred * (gt.median() / pred.median())
        gt, pred = gt.clamp(depth_range[0], depth_range[1]), pred.clamp(depth_range[0], depth_range[1])
        for k, v in _depth_stats(gt, pred).items():
            total[k] = total.get(k, 0) + v
    return {k: v / depth_gt.shape[0] for k, v in total.items()}


def depth_metrics_per_class(depth_gt, depth_pred, semantic_gt, depth_range, use_gt_scale: bool):
    """DepthEvaluator.compute_depth_metrics_per_class (/root/reference/eval/depth.py:72-106)."""
    depth_gt, semantic_gt = depth_gt.unsqueeze(1), semantic_gt.unsqueeze(1)
    out = {}
    for c in torch.unique(semantic_gt):
        if c == 255:
            continue
        gt, pred = depth_gt[semantic_gt == c], depth_pred[semantic_gt == c]
        valid = gt > 0
        if not valid.any():
            continue
        gt, pred = gt[valid], pred[valid]
        if use_gt_scale:
            pred = pred * (gt.median() / pred.median())
        gt, pred = gt.clamp(depth_range[0], depth_range[1]), pred.clamp(depth_range[0], depth_range[1])
        for k, v in _depth_stats(gt, pred).items():
            out[f"{k}_c{int(c)}"] = v
    return out

