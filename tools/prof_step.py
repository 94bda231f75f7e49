import sys, torch
sys.path.insert(0, "/root/repo")
import codeps_b200
from codeps_b200 import synthetic
dev = torch.device("cuda:0")
tb = synthetic.make_preset_batch("cityscapes", 8, seed=1).to(dev)
fn = codeps_b200.ReconstructionLoss(tb.width, tb.height, codeps_b200.SSIMLoss(), 5, dev)
sm = codeps_b200.EdgeAwareSmoothnessLoss()
for it in range(3):
    depth = tb.depth.clone().requires_grad_(True); disp = tb.disp.clone().requires_grad_(True)
    poses = [p.clone().requires_grad_(True) for p in tb.poses]
    loss = 10 * fn(tb.camera_models(), tb.images, depth, poses) + 1e-3 * sm(tb.images[0], disp)
    loss.backward()
torch.cuda.synchronize()
print("ok", float(loss))
