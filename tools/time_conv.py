"""Times the 3x3 decoder convolution (tcgen05 path) alone: python tools/time_conv.py [Cin=360] [Cout=112] [H=56]
MLIIS_TC_DEBUG=32 prints the phase timestamps of CTA 0 (tc_conv3_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from mliis_b200 import native as N
lib = N.lib()
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 360
Cout = int(sys.argv[2]) if len(sys.argv) > 2 else 112
H = int(sys.argv[3]) if len(sys.argv) > 3 else 56
B = 8
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(B,H,H,Cin, device='cuda', generator=g); w = torch.randn(3,3,Cin,Cout, device='cuda', generator=g)*0.02
bias = torch.zeros(Cout, device='cuda'); y = torch.empty(B,H,H,Cout, device='cuda'); wt = torch.empty(2*9*Cin*Cout, device='cuda')
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
quiet = int(os.environ.get('MLIIS_TC_DEBUG', '0')) & 32
for mode in (2,) if quiet else (2,1):
    N.check(lib.mliis_tc_prep_weights(w.data_ptr(), wt.data_ptr(), 9, Cin, Cout, 0, mode, st))
    ts=[]
    for it in range(2 if quiet else 8):
        flush.zero_()
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B,H,H,Cin,Cout,9,1,mode,st)); e1.record(); e1.synchronize()
        if it>=(1 if quiet else 3): ts.append(e0.elapsed_time(e1)*1e3)
    tw=[]
    for it in range(0 if quiet else 6):
        e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); N.check(lib.mliis_tc_conv(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B,H,H,Cin,Cout,9,1,mode,st)); e1.record(); e1.synchronize()
        if it>=2: tw.append(e0.elapsed_time(e1)*1e3)
    fl = 2.0*B*H*H*9*Cin*Cout
    print('Cin=%d Cout=%d H=%d DEBUG=%s mode=%d  cold %.1f us (%.0f TFLOP/s)  warm %.1f us' % (Cin, Cout, H, os.environ.get('MLIIS_TC_DEBUG','0'), mode, np.mean(ts), fl/np.mean(ts)/1e6, np.mean(tw) if tw else -1))
