import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from mliis_b200 import native as N
lib = N.lib()
st = None
def run(x, w, dil, mode):
    B,H,W_,Cin = x.shape; Cout = w.shape[-1]
    wt = torch.empty(2*9*Cin*Cout, device='cuda'); y = torch.full((B,H,W_,Cout), float('nan'), device='cuda')
    N.check(lib.mliis_tc_prep_weights(w.data_ptr(), wt.data_ptr(), 9, Cin, Cout, 0, mode, st))
    N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), None, y.data_ptr(), B,H,W_,Cin,Cout,9,dil,mode,st))
    torch.cuda.synchronize(); return y
for (H, dil) in [(56,1),(14,1),(56,2)]:
    B, C = 1, 32
    # x[b,y,x,c] encodes position: value = y*100 + x (same for all channels) so shifts are readable
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(H), indexing='ij')
    x = (yy*100 + xx).float()[None,:,:,None].expand(B,H,H,C).contiguous().cuda()
    for ky in range(3):
        for kx in range(3):
            w = torch.zeros(3,3,C,C, device='cuda'); w[ky,kx] = torch.eye(C, device='cuda')
            y = run(x, w, dil, 1)
            exp = torch.zeros_like(x)
            dy, dx = (ky-1)*dil, (kx-1)*dil
            ys, xs = slice(max(0,-dy), H-max(0,dy)), slice(max(0,-dx), H-max(0,dx))
            exp[:, ys, xs] = x[:, max(0,dy):H+min(0,dy) or None, max(0,dx):H+min(0,dx) or None][:, :exp[:,ys,xs].shape[1], :exp[:,ys,xs].shape[2]]
            err = (y-exp).abs().max().item()
            # sample a few interior outputs to read the actual source position
            p = [(5,7),(6,7),(min(H-1,9),min(H-1,12))]
            got = [int(round(y[0,a,b,0].item())) for a,b in p]
            want = [int(exp[0,a,b,0].item()) for a,b in p]
            print('H=%d dil=%d tap(%d,%d) maxerr=%.1f got=%s want=%s nan=%d' % (H,dil,ky,kx,err,got,want,torch.isnan(y).sum().item()))
