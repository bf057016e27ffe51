import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location('tm', 'tests/test_gpu_multiclass.py'); tm = importlib.util.module_from_spec(spec); spec.loader.exec_module(tm)
from oracle.efficientlab_oracle import resize_bilinear_ac
from mliis_b200 import native as N
n_classes, size, B = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 64, int(sys.argv[2]) if len(sys.argv) > 2 else 4
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n_ex = int(sys.argv[5]) if len(sys.argv) > 5 else 6
arch, theta, bn, images, masks, cls, dense = tm._problem(n_classes, size, n=n_ex, seed=seed)
eng = tm._engine(arch, theta, bn, size, B, n_classes, gemm_mode=int(sys.argv[3]) if len(sys.argv) > 3 else N.GEMM_FP32)
idx = (np.array([3, 0, 5, 2], np.int32)[:B] if n_ex >= 6 else np.arange(B, dtype=np.int32))
xd, md = torch.from_numpy(images).cuda(), torch.from_numpy(masks).cuda()
eng.set_class_ids(0, torch.from_numpy(cls).cuda())
di = torch.from_numpy(idx).cuda()
z_lo = eng.forward(0, xd, True, index=di)
loss, grads = eng.loss_backward(0, md, B, index=di)
torch.cuda.synchronize()
dz = eng.debug_buffer(0, "head.dlogits_lowres", B).cpu().double().reshape(B, size//4, size//4, n_classes+1)
z = z_lo.cpu().double().requires_grad_(True)
y = torch.from_numpy(dense[idx]).double()
up = resize_bilinear_ac(z.permute(0,3,1,2), size, size).permute(0,2,3,1)
logp = torch.log_softmax(up, -1); ce = -(y*logp).sum(-1).mean()
p = torch.softmax(up, -1).reshape(B, -1); yy = y.reshape(B, -1)
inter = (p*yy).sum(1); den = p.sum(1)+yy.sum(1)-inter
iou = ((inter+1e-7)/(den+1e-7)).mean()
l = ce - torch.log(2*iou/(iou+1))
l.backward()
ref = z.grad
print('loss (no l2) ref', l.item(), 'ce', ce.item(), 'iou', iou.item())
print('dz rel l2', ((dz-ref).norm()/ref.norm()).item(), 'max abs', (dz-ref).abs().max().item(), 'ref max', ref.abs().max().item())
e = (dz-ref).abs()
w = np.unravel_index(e.argmax().item(), e.shape); print('worst at', w, dz[w].item(), ref[w].item())
# per-row errors: borders vs interior
er = e.sum(-1)[0]
ec = e.sum((0,1,2)); print('err by channel: first', ec[:6].tolist(), 'max at', int(ec.argmax()), float(ec.max()), 'cls', cls[idx])
print('err by row y (img0):', [float('%.2e' % v) for v in er.sum(1)])
print('err by col x (img0):', [float('%.2e' % v) for v in er.sum(0)])
from oracle.efficientlab_oracle import EfficientLabOracle
from tests.parity_util import per_param_report, rel_l2
orc = EfficientLabOracle(arch, torch.float64, binary_iou_loss=False)
loss_o, g_o, bn_o, logits_o = orc.loss_and_grad(theta, bn, torch.from_numpy(images[idx]), torch.from_numpy(dense[idx]))
upd = resize_bilinear_ac(z_lo.cpu().double().permute(0,3,1,2), size, size).permute(0,2,3,1)
print('logits max abs err', (upd-logits_o).abs().max().item(), 'logits max', logits_o.abs().max().item())
g = eng.tf_order_vector(grads).cpu().double()
print('loss', loss.item(), loss_o.item(), 'grad rel_l2', rel_l2(g, g_o))
for r in per_param_report(arch, g, g_o, top=10): print(r)
