"""Times the MBConv pointwise convolutions (tc_conv_kernel) alone, n_group task slots per launch:
python tools/time_conv1x1.py [n_group=1].  MLIIS_TC_DEBUG=32 prints the phase timestamps of one CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mliis_b200 import native as N
lib = N.lib()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
st = torch.cuda.current_stream().cuda_stream
flush = bench._Flusher()
quiet = int(os.environ.get('MLIIS_TC_DEBUG', '0')) & 32
if quiet: bench._time_launch.__defaults__ = (1, 1)
B = 8
for (HW, Cin, Cout, fused) in ((112*112, 16, 96, False), (112*112, 96, 24, True), (56*56, 24, 144, False), (56*56, 144, 24, True),
                               (28*28, 40, 240, False), (28*28, 240, 40, True), (14*14, 112, 672, False), (14*14, 672, 112, True)):
    M = B * HW
    ar = bench._Arena(G, dict(x=M*Cin, w=Cin*Cout, wt=2*Cin*Cout, a=Cin, b=Cin, gate=B*Cin, y=M*Cout))
    N.check(lib.mliis_kernel_group(G, ar.stride*4))
    N.check(lib.mliis_tc_prep_weights(ar.p("w"), ar.p("wt"), 1, Cin, Cout, 0, N.GEMM_TF32X3, st))
    if fused:
        fn = lambda: N.check(lib.mliis_tc_project_conv(ar.p("x"), ar.p("wt"), ar.p("a"), ar.p("b"), ar.p("gate"), ar.p("y"), B, HW, Cin, Cout, N.GEMM_TF32X3, st))
    else:
        fn = lambda: N.check(lib.mliis_tc_conv(ar.p("x"), ar.p("wt"), None, ar.p("y"), B, 1, HW, Cin, Cout, 1, 1, N.GEMM_TF32X3, st))
    ms = bench._time_launch(fn, flush)
    torch.cuda.synchronize()
    nb = 4.0 * G * (M*Cin + M*Cout)
    print("%s %4d->%4d M=%6d x%d slots: %7.1f us  %6.0f GB/s (%.2f of 6550)" % ("project" if fused else "expand ", Cin, Cout, M, G, ms*1e3, nb/ms/1e6, nb/ms/1e6/6550.1))
    N.check(lib.mliis_kernel_group(1, 0))
    del ar
