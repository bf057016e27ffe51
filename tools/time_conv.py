import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from mliis_b200 import native as N
lib = N.lib()
B,H,Cin,Cout = 8,56,360,112
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(B,H,H,Cin, device='cuda', generator=g); w = torch.randn(3,3,Cin,Cout, device='cuda', generator=g)*0.02
bias = torch.zeros(Cout, device='cuda'); y = torch.empty(B,H,H,Cout, device='cuda'); wt = torch.empty(2*9*Cin*Cout, device='cuda')
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for mode in (2,1):
    N.check(lib.mliis_tc_prep_weights(w.data_ptr(), wt.data_ptr(), 9, Cin, Cout, 0, mode, st))
    ts=[]
    for it in range(8):
        flush.zero_()
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B,H,H,Cin,Cout,9,1,mode,st)); e1.record(); e1.synchronize()
        if it>=3: ts.append(e0.elapsed_time(e1)*1e3)
    # warm (no flush)
    tw=[]
    for it in range(6):
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B,H,H,Cin,Cout,9,1,mode,st)); e1.record(); e1.synchronize()
        if it>=2: tw.append(e0.elapsed_time(e1)*1e3)
    print('DEBUG=%s mode=%d  cold %.0f us  warm %.0f us' % (os.environ.get('MLIIS_TC_DEBUG','0'), mode, np.mean(ts), np.mean(tw)))
