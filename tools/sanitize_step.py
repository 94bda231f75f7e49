"""Small forward+backward of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, torch
sys.path.insert(0, "/root/repo")
import codeps_b200
from codeps_b200 import synthetic
dev = torch.device("cuda:0")
for (w, h, b, scales) in ((132, 70, 2, 5), (64, 32, 3, 4), (160, 128, 2, 2)):  # the last one has interior tiles
    tb = synthetic.make_batch(b, w, h, (0.8 * w, 0.8 * w, 0.5 * w, 0.5 * h), seed=1, flip_every_other=True).to(dev)
    fn = codeps_b200.ReconstructionLoss(w, h, codeps_b200.SSIMLoss(), scales, dev, noise=("fused" if "--fused" in sys.argv else "torch"))
    sm = codeps_b200.EdgeAwareSmoothnessLoss()
    depth = tb.depth.clone().requires_grad_(True); disp = tb.disp.clone().requires_grad_(True)
    poses = [p.clone().requires_grad_(True) for p in tb.poses]
    loss = 10 * fn(tb.camera_models(), tb.images, depth, poses) + 1e-3 * sm(tb.images[0], disp)
    loss.backward()
    with torch.no_grad():
        fn(tb.camera_models(), tb.images, tb.depth, tb.poses)
    warper = codeps_b200.ImageWarper(w, h, dev)
    d2 = tb.depth.clone().requires_grad_(True); p2 = tb.poses[0].clone().requires_grad_(True)
    out = warper(tb.camera_models(), tb.images[1], d2, p2)
    out.sum().backward()
    x = tb.images[1].clone().requires_grad_(True); y = tb.images[0].clone().requires_grad_(True)
    codeps_b200.SSIMLoss()(x, y).sum().backward()
    torch.cuda.synchronize()
    print("ok", w, h, float(loss))
    # widened rows: object motion, device intrinsics, flow regularisers, heads, c2c warp, depth metrics
    gen = torch.Generator().manual_seed(3)
    motions = [(0.01 * torch.randn(b, 3, h, w, generator=gen)).to(dev).requires_grad_(True) for _ in range(2)]
    k_dev = tb.intrinsics.to(dev)
    lazy = [codeps_b200.CameraModel.from_tensor(w, h, k_dev[i]) for i in range(b)]
    d3 = tb.depth.clone().requires_grad_(True)
    l3 = fn(lazy, tb.images, d3, tb.poses, motions)
    l3 = l3 + codeps_b200.FlowSmoothnessLoss()(tuple(motions)) + codeps_b200.FlowSmoothnessLoss(False)(tuple(motions)) \
        + codeps_b200.FlowSparsityLoss()(tuple(motions))
    l3.backward()
    aa = (0.01 * torch.randn(b, 1, 3, generator=gen)).to(dev).requires_grad_(True)
    tr = (0.1 * torch.randn(b, 1, 3, generator=gen)).to(dev).requires_grad_(True)
    (codeps_b200.transformation_from_parameters(aa, tr, True).sum() + codeps_b200.disp_to_depth(disp).sum()).backward()
    # fused heads entry (disparity + 6-DoF parameters in, gradients w.r.t. them out) and the semantic_mask branch
    aa2 = [(0.01 * torch.randn(b, 1, 3, generator=gen)).to(dev).requires_grad_(True) for _ in range(2)]
    tr2 = [(0.02 * torch.randn(b, 1, 3, generator=gen)).to(dev).requires_grad_(True) for _ in range(2)]
    disp2 = tb.disp.clone().requires_grad_(True)
    l4, _depth, _poses = fn.forward_from_heads(tb.camera_models(), tb.images, disp2, ((aa2[0], tr2[0]), (aa2[1], tr2[1])))
    l4.backward()
    labels = tuple(torch.randint(0, 19, (b, h, w), generator=gen).to(dev) for _ in range(3))
    fn(tb.camera_models(), tb.images, tb.depth, tb.poses, None, labels)
    tgt = torch.zeros(b, 3, h + 6, w + 10, device=dev)
    lbl = torch.randint(0, 9, (b, h, w), generator=gen).to(dev)
    cams_t = [codeps_b200.CameraModel(w + 10, h + 6, 0.7 * w, 0.7 * w, 0.5 * w + 3, 0.5 * h + 2) for _ in range(b)]
    for mode, pad, src in (("bilinear", "zeros", tb.images[0]), ("nearest", "border", lbl)):
        codeps_b200.warp_c2c(tb.camera_models(), cams_t, src, tgt, interp_mode=mode, padding_mode=pad)
    gt = tb.depth[:, 0] * (torch.rand(b, h, w, generator=gen).to(dev) > 0.7)
    ev = codeps_b200.DepthEvaluator(True, (0.1, 80.0), True)
    ev.compute_depth_metrics(gt, tb.depth)
    codeps_b200.DepthEvaluator(False, (0.1, 80.0)).compute_depth_metrics_per_class(gt, tb.depth, lbl % 3)
    torch.cuda.synchronize()
    print("ok widened", w, h, float(l3))
