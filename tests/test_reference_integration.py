"""Row a11 of SURVEY.md section 8: the unchanged CoDEPS code runs on top of the drop-ins.

test_install_against_the_real_checkout (build container only: needs /root/reference) imports the
real checkout, calls codeps_b200.install(), and checks that every rebound class has the
reference's constructor / call signature and public methods, that the constructor calls of
codeps/model_setup.py:63-85 work with the names as that module sees them, and that a real
algos.depth.DepthAlgo is built around the CUDA-backed loss objects.

test_depth_algo_training_call_sequence (GPU) replays DepthAlgo._forward + .training
(algos/depth.py:373-420,466-481) over the drop-ins with random-initialised heads and compares the
losses and the parameter gradients with the same graph evaluated through the oracle on the CPU.
"""
import json
import os
import subprocess
import sys

import pytest
import torch

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "algos")), reason="needs the CoDEPS checkout (build container)")
def test_install_against_the_real_checkout():
    proc = subprocess.run([sys.executable, os.path.join(HERE, "ref_install_check.py"), REFERENCE],
                          capture_output=True, text=True, timeout=600)
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("REPORT ")]
    assert proc.returncode == 0 and lines, proc.stderr[-3000:]
    rep = json.loads(lines[-1][7:])
    assert not rep["errors"], rep["errors"]
    patched = set(rep["patched"])
    for name in ("misc.ImageWarper", "misc.CameraModel", "algos.depth.ReconstructionLoss", "algos.depth.SSIMLoss",
                 "algos.depth.EdgeAwareSmoothnessLoss", "codeps.model_setup.ReconstructionLoss",
                 "codeps.model_setup.DepthEvaluator", "codeps.online_adap.CameraModel", "eval.DepthEvaluator",
                 "datasets.mixup.Mixup.warp_c2c", "models.pose_head.PoseHead.transformation_from_parameters",
                 "models.depth_head.DepthHead.disp_to_depth"):
        assert name in patched, f"{name} was not rebound (patched: {sorted(patched)})"
    for name, entry in rep["signatures"].items():
        assert entry["ours"], f"{name} is still the reference's"
        assert entry["init_equal"], f"{name}.__init__ differs: {entry.get('init_ref')} vs {entry.get('init_ours')}"
        assert entry["extra_init_have_defaults"], f"{name}.__init__ has extra required parameters"
        assert entry["call_equal"], f"{name} call signature differs: {entry.get('call_ref')} vs {entry.get('call_ours')}"
        assert not entry["missing_methods"], f"{name} lacks {entry['missing_methods']}"
    algo = rep["depth_algo"]
    assert algo["is_reference_class"], "DepthAlgo itself must stay the reference's class"
    assert all(m.startswith("codeps_b200") for m in algo["loss_classes"]), algo["loss_classes"]
    assert algo["training_signature"][:6] == ["images", "depth_feats_window", "camera_models", "depth_head",
                                               "body_pose_sflow", "pose_head"]
    assert algo["image_warpers"] == [0, 1, 2, 3, 4] and algo["scaled_width"] == 88
    assert rep["camera_model"]["module"].startswith("codeps_b200")
    assert abs(rep["camera_model"]["scaled_fx"] - 552.5 / 2) < 1e-3
    assert rep["restored"]


@pytest.mark.gpu
def test_depth_algo_training_call_sequence(cuda_device):
    """DepthAlgo._forward / .training over the drop-ins, heads with random weights."""
    sys.path.insert(0, os.path.dirname(HERE))
    import bench_adapt
    import codeps_b200
    from codeps_b200.synthetic import make_batch
    from oracle import photo_oracle as po
    dev = cuda_device
    torch.manual_seed(3)
    w, h, b, scales = 256, 128, 2, 4
    tb = make_batch(b, w, h, (230.0, 232.0, 127.0, 63.0), seed=9, shift_px=2, flip_every_other=True)
    backbone = bench_adapt.Encoder(18).eval()
    nets = bench_adapt.Trainable()
    nets.depth_head = bench_adapt.DepthDecoder(backbone.num_ch_enc)
    nets.eval()  # batch-norm statistics fixed: the two evaluations below must see the same function
    recon_fn = codeps_b200.ReconstructionLoss(w, h, codeps_b200.SSIMLoss(), scales, dev)
    recon_fn.keep_noise = True
    smooth_fn = codeps_b200.EdgeAwareSmoothnessLoss()

    heads_out = {}

    def forward(device, use_cuda_ops):
        """algos/depth.py:373-420 (_forward, no flow head) then :466-481 (training)."""
        bb, net = backbone.to(device), nets.to(device)
        images = tuple(i.to(device) for i in tb.images)
        with torch.no_grad():
            feats = bb(images[0])
        disp = net.depth_head(feats)
        aa0, t0 = net.pose_head(net.pose_encoder(torch.cat([images[1], images[0]], 1)))
        aa1, t1 = net.pose_head(net.pose_encoder(torch.cat([images[0], images[2]], 1)))
        outs = dict(disp=disp, aa0=aa0, t0=t0, aa1=aa1, t1=t1)
        for v in outs.values():
            v.retain_grad()
        heads_out[use_cuda_ops] = outs
        if use_cuda_ops:
            depth = codeps_b200.disp_to_depth(disp)                                      # DepthHead.disp_to_depth
            poses = [codeps_b200.transformation_from_parameters(aa0, t0, invert=True),   # PoseHead.forward
                     codeps_b200.transformation_from_parameters(aa1, t1)]
            cams = [codeps_b200.CameraModel.from_tensor(w, h, k) for k in tb.intrinsics.to(device)]  # online_adap.py:95-100
            recon = recon_fn(cams, images, depth, poses, None)                           # depth.py:474-479
            smooth = smooth_fn(images[0], disp)                                          # depth.py:480
        else:
            depth = po.disp_to_depth(disp)
            poses = [po.transformation_from_parameters(aa0, t0, True), po.transformation_from_parameters(aa1, t1, False)]
            recon = po.reconstruction_loss(tb.intrinsics.numpy(), images, depth, poses, noise, scales)
            smooth = po.smoothness_loss(images[0], disp)
        net.zero_grad(set_to_none=True)
        (10.0 * recon + 0.001 * smooth).backward()
        grads = {n: p.grad.detach().cpu().clone() for n, p in net.named_parameters() if p.grad is not None}
        return float(recon), float(smooth), grads

    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False  # the CPU evaluation below is plain fp32
    try:
        recon_g, smooth_g, grads_g = forward(dev, True)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    noise = [n.cpu() for n in recon_fn.last_noise]
    recon_c, smooth_c, grads_c = forward(torch.device("cpu"), False)
    assert abs(recon_g - recon_c) <= 2e-4 * abs(recon_c), (recon_g, recon_c)   # cuDNN vs CPU convolutions upstream
    assert abs(smooth_g - smooth_c) <= 2e-4 * abs(smooth_c), (smooth_g, smooth_c)
    assert set(grads_g) == set(grads_c) and len(grads_g) > 40
    worst = 0.0
    for name, gc in grads_c.items():
        scale = float(gc.abs().max())
        if scale == 0.0:
            continue
        worst = max(worst, float((grads_g[name] - gc).abs().max()) / scale)
    assert worst <= 2e-3, f"parameter gradients differ by {worst:.2e} of their max-abs"
    # the loss op alone, on identical head outputs: the oracle's gradients w.r.t. (disp, axis-angle, translation)
    # for the values the GPU heads produced
    leaves = {k: v.detach().cpu().double().requires_grad_(True) for k, v in heads_out[True].items()}
    images64 = [i.double() for i in tb.images]
    depth = po.disp_to_depth(leaves["disp"])
    poses = [po.transformation_from_parameters(leaves["aa0"], leaves["t0"], True),
             po.transformation_from_parameters(leaves["aa1"], leaves["t1"], False)]
    recon = po.reconstruction_loss(tb.intrinsics.numpy(), images64, depth, poses, noise, scales,
                                   forced_argmin=[a.cpu() for a in recon_fn.last_argmin])
    (10.0 * recon + 0.001 * po.smoothness_loss(images64[0], leaves["disp"])).backward()
    for k, leaf in leaves.items():
        got = heads_out[True][k].grad.cpu().double()
        dev_k = float((got - leaf.grad).abs().max() / leaf.grad.abs().max())
        limit = 1e-4 if k != "disp" else 5e-3  # dL/d disp: per-pixel, a few pixels sit at discrete switches
        assert dev_k <= limit, f"dL/d {k}: {dev_k:.2e} of max-abs"
    print(f"DepthAlgo.training sequence: recon {recon_g:.6f} vs {recon_c:.6f}, worst parameter-gradient deviation {worst:.2e}")
