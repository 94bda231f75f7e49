"""bench_adapt.py -- BASELINE config 4: the full online-adaptation step around the loss op.

What one step does (the depth branch of /root/reference/codeps/online_adap.py:291-401 +
/root/reference/algos/depth.py:373-420,507-568 at the shapes of cfg/adapt_cityscapes_kitti_360.yaml):

  for key in (source 2 @1024x512, target 1 @1408x384, target_replay 2 @1408x384):
      feats  = frozen ResNet-101 backbone on the three frames of the window (no grad, as the reference
               computes po_depth_feats for every offset; only frame t feeds the depth head)
      depth, disp = DepthHead(feats[t])                  (trained)
      T(t->t-1), T(t->t+1) = PoseHead(ResNet-18 on the 6-channel frame pairs)   (trained)
      recon_k, smooth_k = ReconstructionLoss, EdgeAwareSmoothnessLoss
  loss = 10 * sum n_k recon_k / sum n_k + 0.001 * sum n_k smooth_k / sum n_k;  backward;  Adam(1e-4)

The networks are plain torch (torchvision ResNets + decoders written here with the layer shapes of
models/depth_head.py and models/pose_head.py) with random weights; DDP over the ranks.  The step is
timed three ways: with the codeps_b200 loss, with the reference loss (oracle port = the reference's
ATen op sequence, torch eager on the GPU) and with the loss replaced by a trivial sum, which gives
the loss's share of the step for both.
"""
from __future__ import annotations

import time

import torch
import torch.nn as nn
import torch.nn.functional as F

GROUPS = [("source", "cityscapes", 2, False), ("target", "kitti360_cfg", 1, False), ("target_replay", "kitti360_cfg", 2, True)]
RECON_WEIGHT, SMOOTH_WEIGHT = 10.0, 0.001
NUM_SCALES = 5


class Encoder(nn.Module):
    """torchvision ResNet as a 5-level feature pyramid (models/resnet_encoder.py:82-130)."""

    def __init__(self, layers: int, in_channels: int = 3):
        super().__init__()
        import torchvision
        net = getattr(torchvision.models, f"resnet{layers}")(weights=None)
        if in_channels != 3:
            net.conv1 = nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3, bias=False)
        net.fc = nn.Identity()  # only the feature pyramid is used (no parameters without a gradient under DDP)
        self.net = net
        self.num_ch_enc = [64, 64, 128, 256, 512] if layers <= 34 else [64, 256, 512, 1024, 2048]

    def forward(self, x):
        n = self.net
        f0 = n.relu(n.bn1(n.conv1(x)))
        f1 = n.layer1(n.maxpool(f0))
        f2 = n.layer2(f1)
        f3 = n.layer3(f2)
        f4 = n.layer4(f3)
        return [f0, f1, f2, f3, f4]


class DepthDecoder(nn.Module):
    """Layer shapes of models/depth_head.py:12-47,56-80 (monodepth2 decoder with skips)."""

    def __init__(self, num_ch_enc):
        super().__init__()
        dec = [16, 32, 64, 128, 256]
        self.up0, self.up1 = nn.ModuleList(), nn.ModuleList()
        for i in range(5):
            cin = num_ch_enc[-1] if i == 4 else dec[i + 1]
            self.up0.append(nn.Sequential(nn.Conv2d(cin, dec[i], 3, padding=1), nn.ELU(inplace=True)))
            cin = dec[i] + (num_ch_enc[i - 1] if i > 0 else 0)
            self.up1.append(nn.Sequential(nn.Conv2d(cin, dec[i], 3, padding=1), nn.ELU(inplace=True)))
        self.disp = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dec[0], 1, 3))

    def forward(self, feats):
        x = feats[-1]
        for i in range(4, -1, -1):
            x = self.up0[i](x)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = self.up1[i](x)
        return torch.sigmoid(self.disp(x))


class PoseDecoder(nn.Module):
    """Layer shapes of models/pose_head.py:14-54: squeeze + 3 convs -> mean -> 0.01 * (axis-angle, t)."""

    def __init__(self, num_ch_enc):
        super().__init__()
        self.squeeze = nn.Conv2d(num_ch_enc[-1], 256, 1)
        self.c0 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c1 = nn.Conv2d(256, 256, 3, 1, 1)
        self.c2 = nn.Conv2d(256, 6, 1)

    def forward(self, feats):
        x = F.relu(self.squeeze(feats[-1]))
        x = self.c2(F.relu(self.c1(F.relu(self.c0(x)))))
        x = 0.01 * x.mean(3).mean(2).view(-1, 1, 1, 6)
        return x[..., :3][:, 0], x[..., 3:][:, 0]  # [B,1,3] each, as pose_head.py:47-52 slices them


class Trainable(nn.Module):
    def __init__(self):
        super().__init__()
        self.pose_encoder = Encoder(18, in_channels=6)
        self.pose_head = PoseDecoder(self.pose_encoder.num_ch_enc)
        self.depth_head = DepthDecoder([64, 256, 512, 1024, 2048])

    def forward(self, feats_t, pair_prev, pair_next):
        disp = self.depth_head(feats_t)
        aa0, t0 = self.pose_head(self.pose_encoder(pair_prev))
        aa1, t1 = self.pose_head(self.pose_encoder(pair_next))
        return disp, aa0, t0, aa1, t1


def run(args, rank, world, dev, barrier, max_over_ranks, brief: bool = False):
    import codeps_b200
    from codeps_b200 import synthetic
    from oracle import photo_oracle as po  # reference arm of this benchmark only (torch eager on the GPU)

    torch.manual_seed(7 + rank)
    backbone = Encoder(101).to(dev).eval()
    for prm in backbone.parameters():
        prm.requires_grad_(False)
    model = Trainable().to(dev)
    if world > 1:
        # the module runs once per group before the single backward: no in-place buffer broadcast between them
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], broadcast_buffers=False)
    optim = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)

    data, fns = {}, {}
    for key, preset, n, flip in GROUPS:
        tb = synthetic.make_preset_batch(preset, n, seed=100 * rank + len(data), flip_every_other=flip)
        data[key] = dict(tb=tb.to(dev), n=n, k_dev=tb.intrinsics.to(dev), k_np=tb.intrinsics.numpy(),
                         noise=[x.to(dev) for x in po.draw_noise(n, tb.width, tb.height, NUM_SCALES, seed=3)])
        fns[key] = codeps_b200.ReconstructionLoss(tb.width, tb.height, codeps_b200.SSIMLoss(), NUM_SCALES, dev)
    smooth_fn = codeps_b200.EdgeAwareSmoothnessLoss()
    total = sum(d["n"] for d in data.values())

    def losses(mode, d, key, disp, aa0, t0, aa1, t1):
        tb = d["tb"]
        if mode == "none":  # nets only: something that depends on every output
            return disp.mean() + (aa0.sum() + t0.sum() + aa1.sum() + t1.sum()), disp.mean()
        if mode == "ours":
            # fused entry: disparity and 6-DoF parameters go into the op (no conversion kernels)
            cams = [codeps_b200.CameraModel.from_tensor(tb.width, tb.height, d["k_dev"][i]) for i in range(d["n"])]
            recon, _depth, _poses = fns[key].forward_from_heads(cams, tb.images, disp, ((aa0, t0), (aa1, t1)))
            return recon, smooth_fn(tb.images[0], disp)
        depth = po.disp_to_depth(disp)
        poses = (po.transformation_from_parameters(aa0, t0, True), po.transformation_from_parameters(aa1, t1, False))
        recon = po.reconstruction_loss(d["k_np"], tb.images, depth, poses, d["noise"], NUM_SCALES)
        return recon, po.smoothness_loss(tb.images[0], disp)

    def step(mode):
        optim.zero_grad(set_to_none=True)
        recon_t, smooth_t = 0.0, 0.0
        for key, d in data.items():
            tb = d["tb"]
            with torch.no_grad():  # frozen backbone, all three frames of the window (online_adap.py:325-330)
                feats = [backbone(im) for im in tb.images]
            disp, aa0, t0, aa1, t1 = model(feats[0], torch.cat([tb.images[1], tb.images[0]], 1),
                                           torch.cat([tb.images[0], tb.images[2]], 1))
            recon, smooth = losses(mode, d, key, disp, aa0, t0, aa1, t1)
            recon_t = recon_t + recon * (d["n"] / total)
            smooth_t = smooth_t + smooth * (d["n"] / total)
        loss = RECON_WEIGHT * recon_t + SMOOTH_WEIGHT * smooth_t
        loss.backward()
        optim.step()
        return loss.detach()

    def timed(mode, steps, warmup):
        for _ in range(warmup):
            step(mode)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            last = step(mode)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, (time.perf_counter() - t0) * 1e3 / steps, float(last)

    steps, warmup = (args.steps, max(args.warmup, 3)) if not brief else (max(4, min(args.steps, 12)), max(args.warmup, 3))
    ms_none, _, _ = timed("none", steps, warmup)
    ms_ours, wall_ours, loss_ours = timed("ours", steps, warmup)
    ms_ref, wall_ref, loss_ref = timed("reference", max(3, steps // 2), 2)
    out = {
        "metric": "online-adaptation step (frozen ResNet-101 backbone + depth head + ResNet-18 pose net + loss + Adam), "
                  "frame-triplets/sec", "unit": "triplets/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "value": total * world / (ms_ours * 1e-3), "ms_per_step": ms_ours, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "adapt_step", "per_gpu_triplets": total,
                   "groups": {k: f"{d['n']} @{d['tb'].width}x{d['tb'].height}" for k, d in data.items()},
                   "parallelism": f"ddp{world}", "optimizer": "Adam 1e-4", "cudnn_tf32": torch.backends.cudnn.allow_tf32,
                   "shapes": "cfg/adapt_cityscapes_kitti_360.yaml:16-24,45-49"},
        "step_ms": {"codeps_b200_loss": ms_ours, "reference_loss_torch_eager_cuda": ms_ref, "networks_only": ms_none,
                    "host_wall_ms_codeps_b200": wall_ours, "host_wall_ms_reference": wall_ref},
        "loss_share_of_step": max(ms_ours - ms_none, 0.0) / ms_ours,
        "loss_share_of_step_reference": max(ms_ref - ms_none, 0.0) / ms_ref,
        "loss_ms": {"codeps_b200": ms_ours - ms_none, "reference_torch_eager_cuda": ms_ref - ms_none},
        "step_speedup_vs_reference_loss": ms_ref / ms_ours,
        "loss_value": {"codeps_b200": loss_ours, "reference": loss_ref},
    }
    del model, backbone, optim, data
    torch.cuda.empty_cache()
    return out
