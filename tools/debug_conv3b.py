import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from mliis_b200 import native as N
from oracle.efficientlab_oracle import conv2d_same
lib = N.lib()
def trunc(x): return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
def rn(x): return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def run(x, w, bias, dil, mode):
    B,H,W_,Cin = x.shape; Cout = w.shape[-1]
    wt = torch.empty(2*9*Cin*Cout, device='cuda'); y = torch.full((B,H,W_,Cout), float('nan'), device='cuda')
    N.check(lib.mliis_tc_prep_weights(w.data_ptr(), wt.data_ptr(), 9, Cin, Cout, 0, mode, None))
    N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), bias.data_ptr() if bias is not None else None, y.data_ptr(), B,H,W_,Cin,Cout,9,dil,mode,None))
    torch.cuda.synchronize(); return y
g = torch.Generator().manual_seed(0)
for (B,H,Cin,Cout,dil,use_bias) in [(1,14,32,32,1,0),(1,14,64,32,1,0),(2,14,32,32,1,0),(1,14,40,32,1,0),(1,14,32,112,1,0),(1,14,32,32,1,1),(1,56,32,32,1,0),(1,56,64,32,1,0),(2,56,136,112,2,1),(1,20,40,16,1,1)]:
    x = torch.randn(B,H,H,Cin, generator=g); w = torch.randn(3,3,Cin,Cout, generator=g)*0.1; bias = torch.randn(Cout, generator=g) if use_bias else None
    ref = conv2d_same(trunc(x).double().permute(0,3,1,2), rn(w).double(), dilation=dil, bias=bias.double() if use_bias else None).permute(0,2,3,1)
    y = run(x.cuda(), w.cuda(), bias.cuda() if use_bias else None, dil, 1).cpu().double()
    err = (y-ref).abs()
    bad = (err > 1e-3*ref.abs().max()).float()
    print('B=%d H=%d Cin=%d Cout=%d dil=%d bias=%d: maxerr %.3e (ref max %.2f) bad frac %.4f ; bad by image %s ; bad rows(y) %s ; bad cols(n) first %s' % (
        B,H,Cin,Cout,dil,use_bias, err.max().item(), ref.abs().max().item(), bad.mean().item(),
        bad.mean(dim=(1,2,3)).tolist(), [i for i,v in enumerate(bad.mean(dim=(0,2,3)).tolist()) if v>0][:12], [i for i,v in enumerate(bad.mean(dim=(0,1,2)).tolist()) if v>0][:8]))
