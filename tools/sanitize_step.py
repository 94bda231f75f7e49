"""Small forward+backward of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, torch
sys.path.insert(0, "/root/repo")
import codeps_b200
from codeps_b200 import synthetic
dev = torch.device("cuda:0")
for (w, h, b, scales) in ((132, 70, 2, 5), (64, 32, 3, 4)):
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
