import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from mliis_b200 import native as N
torch.manual_seed(0)
lib = N.lib()
def run(B,H,Cin,Cout,taps,mode,a,g):
    dw = torch.full((taps*Cin, Cout), float('nan'), device='cuda')
    rc = lib.mliis_tc_wgrad(a.data_ptr(), g.data_ptr(), dw.data_ptr(), B,H,H,Cin,Cout,taps,1,mode,None)
    torch.cuda.synchronize()
    return rc, dw
for (B,H,Cin,Cout) in [(1,8,32,16),(1,8,128,32),(2,16,96,16)]:
    M=B*H*H
    a = torch.randn(B,H,H,Cin, device='cuda'); g = torch.randn(B,H,H,Cout, device='cuda')
    ref = a.reshape(M,Cin).double().T @ g.reshape(M,Cout).double()
    for mode in (1,2):
        rc, dw = run(B,H,Cin,Cout,1,mode,a,g)
        d = dw.double()
        print('case',(B,H,Cin,Cout),'mode',mode,'rc',rc,'max|dw|',d.abs().max().item(),'max|ref|',ref.abs().max().item(),
              'nan',torch.isnan(d).sum().item(),'nonzero',(d!=0).sum().item(),'err',(d-ref).abs().max().item())
    # structured probes
    a1 = torch.ones_like(a); g1 = torch.ones_like(g)
    rc, dw = run(B,H,Cin,Cout,1,1,a1,g1); print('  ones x ones -> expect',M,'got uniq',torch.unique(dw).tolist()[:8])
    ar = torch.zeros_like(a); ar[...,:] = torch.arange(Cin, device='cuda').float()
    rc, dw = run(B,H,Cin,Cout,1,1,ar,g1); print('  chan-ramp x ones -> row c should be c*M; got rows', (dw[:,0]/M).tolist()[:40])
    gr = torch.zeros_like(g); gr[...,:] = torch.arange(Cout, device='cuda').float()
    rc, dw = run(B,H,Cin,Cout,1,1,a1,gr); print('  ones x chan-ramp -> col n should be n*M; got cols', (dw[0,:]/M).tolist()[:40])
