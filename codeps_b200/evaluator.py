"""Drop-in ``DepthEvaluator`` (/root/reference/eval/depth.py:7-133) backed by the CUDA metrics
kernels.

``compute_depth_metrics`` runs inside every training step that has ground-truth depth
(/root/reference/algos/depth.py:468-469).  The reference loops over the images in python,
compacts each with a boolean mask (a host synchronisation per image) and sorts twice per image
for the medians; here all images go through one radix-select + one statistics pass without
touching the host.  Same constructor, same method names, same result keys.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import Tensor

from . import ops

STAT_KEYS = ("d_a1", "d_a2", "d_a3", "d_rmse", "d_rmse_log", "d_abs_rel", "d_sq_rel")


class DepthEvaluator:
    """Evaluate depth prediction (same surface as the reference's class)."""

    def __init__(self, use_gt_scale: bool, depth_ranges: Tuple[float, float], use_garg_crop: bool = False):
        self.use_gt_scale = use_gt_scale
        self.depth_ranges = depth_ranges
        self.use_garg_crop = use_garg_crop

    def compute_depth_metrics(self, depth_gt: Tensor, depth_pred: Tensor) -> Dict[str, Tensor]:
        """Batch mean of the per-image metrics over the pixels with ground truth (eval/depth.py:21-70).
        An image without any ground-truth pixel yields NaN (the reference raises in ``median``)."""
        if depth_gt.dim() == 3:
            depth_gt = depth_gt.unsqueeze(1)  # B, H, W -> B, 1, H, W
        out = ops.depth_metrics(depth_gt, depth_pred, self.depth_ranges[0], self.depth_ranges[1],
                                self.use_gt_scale, self.use_garg_crop)
        return {key: out[i] for i, key in enumerate(STAT_KEYS)}

    def compute_depth_metrics_per_class(self, depth_gt: Tensor, depth_pred: Tensor,
                                        semantic_gt: Tensor) -> Dict[str, Tensor]:
        """Metrics of the whole batch restricted to each semantic class (eval/depth.py:72-106).
        The set of keys depends on the data, so the class list and the "class has ground truth"
        flags are read back once (validation only, as in the reference)."""
        depth_gt = depth_gt.unsqueeze(1)
        semantic_gt = semantic_gt.unsqueeze(1).to(torch.int64)
        classes = [int(c) for c in torch.unique(semantic_gt).tolist() if int(c) != 255]
        outs = [ops.depth_metrics(depth_gt, depth_pred, self.depth_ranges[0], self.depth_ranges[1],
                                  self.use_gt_scale, False, labels=semantic_gt, class_id=c) for c in classes]
        depth_stats = {}
        if not outs:
            return depth_stats
        stacked = torch.stack(outs)
        has_gt = (stacked[:, 7] > 0).tolist()
        for row, c, ok in zip(stacked, classes, has_gt):
            if not ok:
                continue
            for i, key in enumerate(STAT_KEYS):
                depth_stats[f"{key}_c{c}"] = row[i]
        return depth_stats

    @staticmethod
    def _compute_depth_stats(gt: Tensor, pred: Tensor) -> Dict[str, Tensor]:
        """Error metrics of two flat tensors of valid depths (eval/depth.py:108-133)."""
        n = gt.numel()
        out = ops.depth_metrics(gt.reshape(1, 1, 1, n), pred.reshape(1, 1, 1, n), float("-inf"), float("inf"),
                                False, False)
        return {key: out[i] for i, key in enumerate(STAT_KEYS)}
