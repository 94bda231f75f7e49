"""Pinhole camera container used at the drop-in boundary.

Mirrors the public surface of the reference's ``misc.CameraModel``
(/root/reference/misc/camera_model.py:9-71): same constructor, attributes and
method names, so callers such as ``CodepsNet.forward``
(/root/reference/codeps/online_adap.py:95-100) keep working.  Everything here is
host-side Python; the per-pixel math lives in the CUDA kernels, which receive
the four intrinsics by value (kernel parameter space = constant bank).

``from_tensor`` of a CUDA tensor is lazy: the object keeps the device tensor and reads it back
only when somebody asks for ``.intrinsics`` on the host.  ``ReconstructionLoss`` hands the device
values straight to its kernels, so the per-sample host synchronisation of the reference
(/root/reference/misc/camera_model.py:27) disappears from the training step.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Sequence, Tuple

import torch

_KEYS = ("fx", "fy", "cx", "cy")


class CameraModel:
    """fx, fy, cx, cy plus the image size they refer to."""

    def __init__(self, img_width, img_height, fx, fy, cx, cy):
        for name, value, strict in (("img_width", img_width, True), ("img_height", img_height, True),
                                    ("fx", fx, False), ("fy", fy, False), ("cx", cx, False),
                                    ("cy", cy, False)):
            # same guards as camera_model.py:12-17 (sizes strictly positive, intrinsics >= 0)
            if strict:
                assert value > 0, f"{name} <= 0 is not allowed"
            else:
                assert value >= 0, f"{name} < 0 is not allowed"
        self.image_size = {"width": img_width, "height": img_height}
        self._intrinsics = OrderedDict(zip(_KEYS, (fx, fy, cx, cy)))
        self._device_intrinsics = None  # CUDA tensor [4] when built lazily by from_tensor

    @property
    def intrinsics(self) -> OrderedDict:
        """fx, fy, cx, cy on the host (read back from the device on first use if built lazily)."""
        if self._intrinsics is None:
            k = self._device_intrinsics.cpu().numpy()
            for name, value in zip(_KEYS, k):
                assert value >= 0, f"{name} < 0 is not allowed"
            self._intrinsics = OrderedDict(zip(_KEYS, (k[0], k[1], k[2], k[3])))
        return self._intrinsics

    @intrinsics.setter
    def intrinsics(self, value) -> None:
        self._intrinsics = value
        self._device_intrinsics = None

    @property
    def device_intrinsics(self):
        """The CUDA tensor [4] this model was built from (None for host-built models)."""
        return self._device_intrinsics

    # ------------------------------------------------------------------ conversions
    def to_tensor(self) -> torch.Tensor:
        return torch.Tensor([self.intrinsics[k] for k in _KEYS])

    @classmethod
    def from_tensor(cls, img_width: int, img_height: int, intrinsics: torch.Tensor) -> "CameraModel":
        # camera_model.py:26-29.  The reference reads the four values back right here (one D2H
        # sync per sample); for a CUDA fp32 tensor that read-back is deferred until the host values
        # are actually asked for.  Values stay numpy float32 scalars so that later scaling happens in
        # the same precision as in the reference.
        t = intrinsics.detach()
        if t.is_cuda and t.dtype == torch.float32 and t.dim() == 1 and t.numel() == 4:
            assert img_width > 0, "img_width <= 0 is not allowed"
            assert img_height > 0, "img_height <= 0 is not allowed"
            cam = cls.__new__(cls)
            cam.image_size = {"width": img_width, "height": img_height}
            cam._intrinsics = None
            cam._device_intrinsics = t
            return cam
        k = t.cpu().numpy()
        return cls(img_width, img_height, k[0], k[1], k[2], k[3])

    def as_tuple(self) -> Tuple[float, float, float, float]:
        """(fx, fy, cx, cy) as python floats (additive helper, not in the reference)."""
        return tuple(float(self.intrinsics[k]) for k in _KEYS)

    # ------------------------------------------------------------------ rescaling
    def get_scaled_model(self, scale_u, scale_v) -> "CameraModel":
        # camera_model.py:31-34
        f = self.intrinsics
        return CameraModel(self.image_size["width"] * scale_u, self.image_size["height"] * scale_v,
                           f["fx"] * scale_u, f["fy"] * scale_v, f["cx"] * scale_u,
                           f["cy"] * scale_v)

    def get_scaled_model_image_size(self, width, height) -> "CameraModel":
        # camera_model.py:36-41 -- the scale is a python float ratio of the two sizes.
        su = width / self.image_size["width"]
        sv = height / self.image_size["height"]
        f = self.intrinsics
        return CameraModel(width, height, f["fx"] * su, f["fy"] * sv, f["cx"] * su, f["cy"] * sv)

    # ------------------------------------------------------------------ projection helpers
    def get_image_point(self, x3d, y3d, z3d):
        """3-D point(s) -> pixel coordinates (camera_model.py:43-50)."""
        f = self.intrinsics
        return (x3d / z3d) * f["fx"] + f["cx"], (y3d / z3d) * f["fy"] + f["cy"]

    def get_viewing_ray(self, u2d, v2d):
        """Pixel coordinates -> unit viewing ray (camera_model.py:52-71)."""
        f = self.intrinsics
        rx = (u2d - f["cx"]) / f["fx"]
        ry = (v2d - f["cy"]) / f["fy"]
        norm = torch.sqrt(rx**2 + ry**2 + 1.0**2)
        return rx / norm, ry / norm, 1.0 / norm

    def __repr__(self) -> str:
        w, h = self.image_size["width"], self.image_size["height"]
        fx, fy, cx, cy = (self.intrinsics[k] for k in _KEYS)
        return f"CameraModel({w}x{h}, fx={fx}, fy={fy}, cx={cx}, cy={cy})"


def level_intrinsics(camera_models: Sequence[CameraModel], width: int, height: int):
    """Per-sample (fx, fy, cx, cy) rescaled to ``width`` x ``height`` exactly the way
    ReconstructionLoss does it (/root/reference/algos/depth.py:273-276), rounded to fp32
    the way the reference rounds them when they meet an fp32 tensor."""
    import numpy as np
    out = np.empty((len(camera_models), 4), dtype=np.float32)
    for i, cam in enumerate(camera_models):
        scaled = cam.get_scaled_model_image_size(width, height)
        for j, k in enumerate(_KEYS):
            out[i, j] = np.float32(scaled.intrinsics[k])
    return out
