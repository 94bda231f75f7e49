import sys, torch
sys.path.insert(0, "/root/repo")
import codeps_b200
from codeps_b200 import synthetic, _native
dev = torch.device("cuda:0")
tb = synthetic.make_preset_batch("cityscapes", 8, seed=1).to(dev)
w, h = tb.width, tb.height
warper = codeps_b200.ImageWarper(w, h, dev)
cams = tb.camera_models()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
with torch.no_grad():
    t_fwd = timeit(lambda: warper(cams, tb.images[1], tb.depth, tb.poses[0]))
depth = tb.depth.clone().requires_grad_(True); pose = tb.poses[0].clone().requires_grad_(True)
out = warper(cams, tb.images[1], depth, pose)
go = torch.ones_like(out)
t_bwd = timeit(lambda: torch.autograd.grad(out, [depth, pose], go, retain_graph=True))
print(f"warp fwd (1 source, 3 ch, 4.19 Mpx): {t_fwd:.1f} us ; warp bwd (taps+deriv gather+adjoint+dT reduce): {t_bwd:.1f} us")
x = tb.images[1]; y = tb.images[0]
with torch.no_grad():
    t_ssim = timeit(lambda: codeps_b200.SSIMLoss()(x, y))
print(f"standalone ssim fwd (24 planes x 0.52 Mpx): {t_ssim:.1f} us")
