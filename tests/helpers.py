"""Shared helpers for the parity tests: golden-fixture loading and tolerance checks."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ("city_near", "kitti_odd", "city_wide")

# Tolerances stated by BASELINE.json's north_star.
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4  # relative to the reference gradient's max-abs
TIE_GAP = 1e-6  # argmin must match wherever the fp64 reference's top-2 gap exceeds this


class Golden:
    """One tests/golden/<name>.npz produced by oracle/make_golden.py from the reference."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        self.width = int(self.z["width"])
        self.height = int(self.z["height"])
        self.num_scales = int(self.z["num_scales"])

    def t(self, key, device="cpu", dtype=None):
        x = torch.from_numpy(np.ascontiguousarray(self.z[key]))
        if dtype is not None:
            x = x.to(dtype)
        return x.to(device)

    def inputs(self, device="cpu"):
        images = tuple(self.t(k, device) for k in ("tgt", "prev", "next"))
        poses = tuple(self.t(k, device) for k in ("pose0", "pose1"))
        noise = [self.t(f"noise{s}", device) for s in range(self.num_scales)]
        return dict(images=images, depth=self.t("depth", device), disp=self.t("disp", device),
                    poses=poses, noise=noise, intrinsics=self.z["intrinsics"])


def rel_err(got, want):
    got = torch.as_tensor(got, dtype=torch.float64).cpu()
    want = torch.as_tensor(want, dtype=torch.float64).cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def assert_loss_close(got, want, what, rtol=LOSS_RTOL):
    got, want = float(got), float(want)
    assert abs(got - want) <= rtol * abs(want), f"{what}: {got!r} vs {want!r} (rtol {rtol})"


def assert_grad_close(got, want, what, rtol=GRAD_RTOL):
    err = rel_err(got, want)
    assert err <= rtol, f"{what}: max-abs-normalised error {err:.3e} > {rtol}"


def assert_argmin_matches(got, golden: Golden, level: int, what=""):
    """Bit-exact wherever the fp64 reference is not within TIE_GAP of a tie."""
    got = torch.as_tensor(got).cpu().to(torch.uint8).numpy()
    ref64 = golden.z[f"ref64_argmin{level}"]
    decided = golden.z[f"ref64_gap{level}"] > TIE_GAP
    bad = (got != ref64) & decided
    assert not bad.any(), (f"{what} level {level}: {int(bad.sum())} argmin mismatches away from ties "
                           f"(of {bad.size}; {int((~decided).sum())} near-tie pixels excluded)")
    return int((got != ref64).sum())
