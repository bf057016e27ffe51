"""unit checks of the tensor-core kernels on the multi-class head shapes (112 <-> 1004 channels, M = 2*16*16)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mliis_b200 import native as N
lib = N.lib()
B, H = 2, 16
g = torch.Generator().manual_seed(0)
def rel(a, b): return ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
for Cin, Cout in ((112, 1004), (1004, 112), (112, 12), (12, 112)):
    x = torch.randn(B, H, H, Cin, generator=g); w = torch.randn(1, 1, Cin, Cout, generator=g) * 0.1
    bias = torch.randn(Cout, generator=g)
    ref = x.double().reshape(-1, Cin) @ w.double().reshape(Cin, Cout) + bias.double()
    xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
    wt = torch.empty(2 * Cin * Cout, device='cuda'); y = torch.full((B * H * H, Cout), float('nan'), device='cuda')
    N.check(lib.mliis_tc_prep_weights(wd.data_ptr(), wt.data_ptr(), 1, Cin, Cout, 0, 2, None))
    N.check(lib.mliis_tc_conv(xd.data_ptr(), wt.data_ptr(), bd.data_ptr(), y.data_ptr(), B, H, H, Cin, Cout, 1, 1, 2, None))
    torch.cuda.synchronize()
    print('fwd  %4d -> %4d  rel err %.2e  nan %d' % (Cin, Cout, rel(torch.nan_to_num(y, nan=1e9), ref), int(torch.isnan(y).sum())))
    # wgrad: a [.., Cin], g [.., Cout] -> dW [Cin, Cout]   (N = Cout must be <= 256)
    if Cout <= 256:
        gr = torch.randn(B, H, H, Cout, generator=g)
        refw = x.double().reshape(-1, Cin).T @ gr.double().reshape(-1, Cout)
        dw = torch.full((Cin, Cout), float('nan'), device='cuda')
        N.check(lib.mliis_tc_wgrad(xd.data_ptr(), gr.cuda().data_ptr(), dw.data_ptr(), B, H, H, Cin, Cout, 1, 1, 2, None))
        torch.cuda.synchronize()
        print('wgrad C=%4d N=%4d rel err %.2e  nan %d' % (Cin, Cout, rel(torch.nan_to_num(dw, nan=1e9), refw), int(torch.isnan(dw).sum())))
