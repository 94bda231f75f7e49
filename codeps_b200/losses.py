"""Drop-in loss classes of the self-supervised depth path, backed by the fused CUDA kernels.

Names, constructor arguments and call signatures follow /root/reference/algos/depth.py:58-326
(``EdgeAwareSmoothnessLoss``, ``SSIMLoss``, ``ReconstructionLoss``; ``:15-52`` for the two flow
regularisers) so that
``codeps.model_setup.gen_models`` (/root/reference/codeps/model_setup.py:63-85) and ``DepthAlgo``
(/root/reference/algos/depth.py:474-481) use them unchanged.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import ops
from .camera import CameraModel
from .warper import ImageWarper


class FlowSmoothnessLoss:
    """Smoothness of the object-motion maps, mean sqrt(dx^2 + dy^2 + 1e-7) of the (wrap-around)
    backward differences, averaged over the maps (/root/reference/algos/depth.py:15-34)."""

    def __init__(self, wrap_around: bool = True):
        self.wrap_around = wrap_around

    def __call__(self, flow_maps: Tuple[Tensor, ...]) -> Tensor:
        return ops.flow_smoothness_loss(flow_maps, self.wrap_around)


class FlowSparsityLoss:
    """Sparsity of the object-motion maps, mean 2 m sqrt(|f| / (m + 1e-7) + 1) with the detached
    spatial mean m of |f| (/root/reference/algos/depth.py:37-52)."""

    def __call__(self, flow_maps: Tuple[Tensor, ...]) -> Tensor:
        return ops.flow_sparsity_loss(flow_maps)


class EdgeAwareSmoothnessLoss:
    """Edge-aware smoothness of the mean-normalised disparity
    (/root/reference/algos/depth.py:58-107): two reduction kernels forward, one kernel backward."""

    def __init__(self):
        pass

    def __call__(self, target_image: Tensor, disparity_map: Tensor) -> Tensor:
        return ops.smoothness_loss(target_image, disparity_map)


class SSIMLoss:
    """3x3 SSIM loss map, clamp((1 - SSIM) / 2, 0, 1) (/root/reference/algos/depth.py:110-155)."""

    def __init__(self, window_size: int = 3):
        if window_size != 3:
            raise NotImplementedError("only the 3x3 window the reference uses is implemented")
        self.window_size = window_size
        self.c1 = .01**2
        self.c2 = .03**2

    def __call__(self, src_img: Tensor, target_img: Tensor) -> Tensor:
        return ops.ssim_loss_map(src_img, target_img)


class ReconstructionLoss:
    """Multi-scale photometric reprojection loss with min-reprojection and identity auto-mask
    (/root/reference/algos/depth.py:176-326).

    One pyramid kernel, one fused tile kernel covering every level (warp of both source frames,
    SSIM + L1 for the two reprojections and the two identity candidates, tie-break noise, min /
    argmin, and -- when a gradient is required -- the complete backward to per-level depth and to
    both poses), and one fixed-order reduction kernel.  ``loss.backward()`` then costs a single
    kernel that assembles dL/d depth from the per-level gradients.

    ``object_motion_maps`` (two [B,3,H,W] maps, depth.py:296-303 -> image_warper.py:133-134) are
    resized with the images, added to the transformed points inside the tile kernel, and receive
    their gradient; ``semantic_mask`` (depth.py:284-292, no caller in the reference) is served by a
    composition of the stand-alone nearest-warp and SSIM kernels (``_semantic_loss``).

    Additive surface (not in the reference, which discards the argmin at depth.py:323):
    ``last_argmin`` -- per level a uint8 [B,H_s,W_s] map, 0/1 = reprojection from t-1/t+1 won,
    2/3 = an identity candidate won, i.e. the pixel is auto-masked.

    ``noise``: "fused" (default) draws the tie-break noise of depth.py:316-318 inside the tile kernel
    from a counter-based generator (hash + Box-Muller: i.i.d. N(0, 1) per pixel and identity candidate;
    no generator launches, no noise tensors in HBM).  Its seed is a device counter that starts at
    ``seed`` (default: ``torch.initial_seed()`` when the object is built, so ``torch.manual_seed``
    controls it) and advances by one per call -- also per replay of a CUDA graph that captured the
    call.  "torch" draws ``torch.randn(B,2,H_s,W_s)`` per level exactly like the reference, so a seeded
    run consumes torch's random stream the same way the reference does (five more launches and
    16 bytes of traffic per level-pixel: 7 % of a step at 1024x512).  ``keep_noise`` keeps the draws
    of the last call in ``last_noise`` in either mode.
    """

    def __init__(self, ref_img_width, ref_img_height, ssim: SSIMLoss, num_scales: int, device: torch.device,
                 alpha: float = .85, noise: str = "fused", seed: Optional[int] = None):
        if noise not in ("torch", "fused"):
            raise ValueError("noise must be 'torch' or 'fused'")
        self.ssim = ssim
        self.device = device
        self.num_scales = num_scales
        self.alpha = alpha
        self.noise = noise
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self._calls = 0
        self._seed_counters = {}  # device index -> one-element int64 tensor (noise="fused")
        self.image_warpers = {}
        self.scaled_width = {}
        self.scaled_height = {}
        for i in range(self.num_scales):
            self.scaled_width[i] = ref_img_width // 2**i
            self.scaled_height[i] = ref_img_height // 2**i
            self.image_warpers[i] = ImageWarper(self.scaled_width[i], self.scaled_height[i], device)
        self.last_argmin: List[Tensor] = []
        self.keep_noise = False  # debugging / tests: keep the tie-break draws of the last call in last_noise
        self.last_noise: Optional[List[Tensor]] = None
        self._k_cache = {}

    def _device_intrinsics(self, camera_models: List[CameraModel], device) -> Optional[Tensor]:
        """[B,4] CUDA tensor of full-resolution intrinsics if every camera model was built lazily
        from a CUDA tensor (CameraModel.from_tensor) for this loss's image size, else None.  Rows of
        one contiguous [B,4] tensor (the usual in_data["camera_model"][i]) are used in place."""
        rows = [getattr(cam, "device_intrinsics", None) for cam in camera_models]
        if not rows or any(r is None or r.device != device for r in rows):
            return None
        if any(cam.image_size["width"] != self.scaled_width[0] or cam.image_size["height"] != self.scaled_height[0]
               for cam in camera_models):
            return None
        base = rows[0]
        step = 4 * base.element_size()
        if all(r.is_contiguous() and r.data_ptr() == base.data_ptr() + i * step and
               r.untyped_storage().data_ptr() == base.untyped_storage().data_ptr() for i, r in enumerate(rows)):
            return torch.as_strided(base, (len(rows), 4), (4, 1))
        return torch.stack(rows)

    def _level_intrinsics(self, camera_models: List[CameraModel]) -> np.ndarray:
        """[num_scales, B, 4]: CameraModel.get_scaled_model_image_size per level, as at
        depth.py:273-276 (cached per distinct camera)."""
        out = np.empty((self.num_scales, len(camera_models), 4), dtype=np.float32)
        for b, cam in enumerate(camera_models):
            key = (cam.image_size["width"], cam.image_size["height"]) + tuple(cam.intrinsics.values())
            rows = self._k_cache.get(key)
            if rows is None:
                rows = np.empty((self.num_scales, 4), dtype=np.float32)
                for s in range(self.num_scales):
                    scaled = cam.get_scaled_model_image_size(self.scaled_width[s], self.scaled_height[s])
                    rows[s] = [np.float32(v) for v in scaled.intrinsics.values()]
                if len(self._k_cache) < 4096:
                    self._k_cache[key] = rows
            out[:, b] = rows
        return out

    def __call__(self, camera_models: List[CameraModel], images: Tuple[Tensor, Tensor, Tensor],
                 depth_map: Tensor, poses: Tuple[Tensor, Tensor],
                 object_motion_maps: Optional[Tuple[Tensor, Tensor]] = None,
                 semantic_mask: Optional[Tuple[Tensor, Tensor, Tensor]] = None) -> Tensor:
        assert len(camera_models) == images[0].shape[0], "Batch size of camera model does not match"
        if semantic_mask is not None:
            return self._semantic_loss(camera_models, depth_map, poses, semantic_mask)
        noise, noise_event, intrinsics = self._prepare(camera_models, images, depth_map)
        loss, self.last_argmin = ops.photometric_loss(
            intrinsics, images, depth_map, poses, noise, self.num_scales,
            self.alpha, seed=self._seed_for(depth_map), motions=object_motion_maps, noise_event=noise_event)
        self._after_call(images[0])
        return loss

    def _semantic_loss(self, camera_models, depth_map, poses, semantic_mask) -> Tensor:
        """The ``semantic_mask`` branch (/root/reference/algos/depth.py:284-292,307-308): label maps
        [B,H,W] of frames t, t-1, t+1; per level nearest resize, nearest-neighbour warp of the two
        neighbouring maps (``cdp_warp_image_fwd``), SSIM (``cdp_ssim_fwd``) + L1 against the target
        map, mean over both candidates and all pixels -- no min-reprojection, no auto-mask.  No
        caller in the reference takes it; a nearest-neighbour warp has a zero coordinate gradient,
        so the value carries no gradient to depth or pose (as in the reference)."""
        import torch.nn.functional as F
        if not depth_map.is_cuda:
            raise RuntimeError("codeps_b200 is CUDA only: depth_map is on the CPU (there is no CPU fallback)")
        loss = torch.zeros((), dtype=torch.float32, device=depth_map.device)
        for s in range(self.num_scales):
            size = (self.scaled_height[s], self.scaled_width[s])
            cams = [cam.get_scaled_model_image_size(self.scaled_width[s], self.scaled_height[s]) for cam in camera_models]
            depth_s = F.interpolate(depth_map, size, mode="bilinear", align_corners=False)
            target = F.interpolate(semantic_mask[0].unsqueeze(1).float(), size, mode="nearest")
            terms = []
            for i, frame in enumerate(semantic_mask[1:]):
                frame_s = F.interpolate(frame.unsqueeze(1).float(), size, mode="nearest")
                pred = self.image_warpers[s](cams, frame_s, depth_s, poses[i], interp_mode="nearest")
                l1 = torch.abs(pred - target).mean(1, True)
                terms.append(self.alpha * self.ssim(pred, target).mean(1, True) + (1 - self.alpha) * l1)
            loss = loss + torch.cat(terms, 1).mean() / (2 ** s)
        return loss / self.num_scales

    def forward_from_heads(self, camera_models: List[CameraModel], images: Tuple[Tensor, Tensor, Tensor],
                           disparity_map: Tensor, pose_parameters: Tuple[Tuple[Tensor, Tensor], Tuple[Tensor, Tensor]],
                           object_motion_maps: Optional[Tuple[Tensor, Tensor]] = None,
                           min_depth: float = 0.1, max_depth: float = 100.0):
        """Additive entry point: the loss on what the heads emit.

        ``disparity_map`` is the depth head's sigmoid output (``DepthHead.disp_to_depth``,
        /root/reference/models/depth_head.py:49-54, runs inside the op) and ``pose_parameters`` the
        pose head's ((axisangle, translation) for t -> t-1, (axisangle, translation) for t -> t+1)
        (``PoseHead.transformation_from_parameters``, /root/reference/models/pose_head.py:56-77,
        runs inside the op; the first pair is inverted as at /root/reference/algos/depth.py:404-407).
        Returns ``(loss, depth_map, (T0, T1))``; gradients go to the disparity and the 6-DoF
        parameters directly, no conversion kernels or autograd nodes of their own in the step."""
        assert len(camera_models) == images[0].shape[0], "Batch size of camera model does not match"
        noise, noise_event, intrinsics = self._prepare(camera_models, images, disparity_map)
        loss, self.last_argmin, depth, transformations = ops.photometric_loss_from_heads(
            intrinsics, images, disparity_map, pose_parameters, noise, self.num_scales, self.alpha,
            seed=self._seed_for(disparity_map), min_depth=min_depth, max_depth=max_depth, motions=object_motion_maps,
            noise_event=noise_event)
        self._after_call(images[0])
        return loss, depth, transformations

    def _prepare(self, camera_models, images, depth_map):
        """Shape check, tie-break noise (drawn on a side stream) and intrinsics for one call."""
        h, w = images[0].shape[2], images[0].shape[3]
        if (w, h) != (self.scaled_width[0], self.scaled_height[0]):
            raise ValueError(f"images are {w}x{h} but this loss was built for "
                             f"{self.scaled_width[0]}x{self.scaled_height[0]}")
        b = images[0].shape[0]
        noise, noise_event = None, None
        if self.noise == "torch" and not depth_map.is_cuda:
            noise = [torch.randn((b, 2, self.scaled_height[s], self.scaled_width[s]), device=depth_map.device)
                     for s in range(self.num_scales)]  # rejected by the op with the "CUDA only" error
        elif self.noise == "torch":
            # drawn on a side stream so that the five randn kernels overlap the pyramid kernel (both
            # are independent and memory-bound); same generator, same call order, same values
            device = depth_map.device
            current = torch.cuda.current_stream(device)
            side = self._side_stream(device)
            side.wait_stream(current)
            with torch.cuda.stream(side):
                noise = [torch.randn((b, 2, self.scaled_height[s], self.scaled_width[s]), device=device)
                         for s in range(self.num_scales)]
                noise_event = torch.cuda.Event()
                noise_event.record(side)
            for n in noise:
                n.record_stream(current)
        self._calls += 1
        self.last_noise = noise if self.keep_noise else None
        intrinsics = self._device_intrinsics(camera_models, depth_map.device)
        if intrinsics is None:
            intrinsics = self._level_intrinsics(camera_models)
        return noise, noise_event, intrinsics

    def _seed_for(self, like: Tensor):
        """noise="fused": the device seed counter of ``like``'s device (created on first use); else the
        host seed (unused by the kernels when noise tensors are passed)."""
        if self.noise != "fused" or not like.is_cuda:
            return self.seed
        key = like.device.index if like.device.index is not None else torch.cuda.current_device()
        counter = self._seed_counters.get(key)
        if counter is None:
            counter = self._seed_counters[key] = torch.full((1,), self.seed, dtype=torch.int64, device=like.device)
        return counter

    def noise_seed_state(self, device=None) -> int:
        """Seed the next noise="fused" call on ``device`` will use (reads the device counter: synchronises)."""
        device = torch.device(self.device if device is None else device)
        key = device.index if device.index is not None else torch.cuda.current_device()
        counter = self._seed_counters.get(key)
        return self.seed if counter is None else int(counter.item())

    def reset_noise_seed(self, seed: Optional[int] = None) -> None:
        """Restart the noise="fused" generator at ``seed`` (default: the construction seed) on every device."""
        self.seed = self.seed if seed is None else int(seed) & 0x7FFFFFFFFFFFFFFF
        for counter in self._seed_counters.values():
            counter.fill_(self.seed)

    def _after_call(self, like: Tensor):
        """keep_noise with the built-in generator: materialise the draws the call just used (reads the
        device counter back: a host synchronisation, debugging / tests only)."""
        if self.keep_noise and self.noise == "fused" and like.is_cuda and not torch.cuda.is_current_stream_capturing():
            used = int(self._seed_for(like).item()) - 1
            b = like.shape[0]
            self.last_noise = ops.tiebreak_noise(b, self.scaled_height[0], self.scaled_width[0], self.num_scales, used,
                                                 like.device)

    _side_streams = {}

    @classmethod
    def _side_stream(cls, device) -> "torch.cuda.Stream":
        """One auxiliary stream per device for the tie-break noise (shared by all loss objects)."""
        key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        stream = cls._side_streams.get(key)
        if stream is None:
            stream = cls._side_streams[key] = torch.cuda.Stream(device=key)
        return stream

    def auto_mask(self, level: int = 0) -> Tensor:
        """Boolean [B,H_s,W_s]: True where the last call auto-masked the pixel (identity won)."""
        return self.last_argmin[level] >= 2
