"""Run a handful of representative SlowFast-R50 (B=8) convolution layers once each, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from pytorchvideo_b200 import _lib as L
from pytorchvideo_b200.engine.plan import Plan

LAYERS = [
    # name, (N,Ci,T,H,W), Co, k, s, p, residual
    ("fast_stem_5x7x7", (8, 3, 32, 224, 224), 8, (5, 7, 7), (1, 2, 2), (2, 3, 3), False),
    ("slow_stem_1x7x7", (8, 3, 8, 224, 224), 64, (1, 7, 7), (1, 2, 2), (0, 3, 3), False),
    ("res2_conv_c_pw64to256_res", (8, 64, 8, 56, 56), 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("res2_conv_b_1x3x3_64", (8, 64, 8, 56, 56), 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("res4_conv_a_3x1x1_1024to256", (8, 1024, 8, 14, 14), 256, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),
    ("res4_conv_b_1x3x3_256", (8, 256, 8, 14, 14), 256, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("res5_conv_c_pw512to2048_res", (8, 512, 8, 7, 7), 2048, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("fast_res2_conv_b_1x3x3_8", (8, 8, 32, 56, 56), 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("fast_res2_conv_a_3x1x1_32to8", (8, 32, 32, 56, 56), 8, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),
]

def main():
    only = sys.argv[1:] 
    dev = torch.device("cuda:0")
    for name, xs, co, k, s, p, use_res in LAYERS:
        if only and not any(o in name for o in only):
            continue
        plan = Plan(dev, L.PV_F16, True)
        x = torch.randn(xs, device=dev)
        xr = plan.emit_input_ncdhw(x, xs[1], 4 if xs[1] <= 4 else xs[1])
        w = torch.randn(co, xs[1], *k) * 0.05
        bn = nn.BatchNorm3d(co).eval()
        # shape of output for the residual
        To = (xs[2] + 2 * p[0] - k[0]) // s[0] + 1; Ho = (xs[3] + 2 * p[1] - k[1]) // s[1] + 1; Wo = (xs[4] + 2 * p[2] - k[2]) // s[2] + 1
        rr = None
        if use_res:
            r = torch.randn(xs[0], co, To, Ho, Wo, device=dev)
            rr = plan.emit_input_ncdhw(r, co, co)
        y = plan.emit_conv(xr, w, None, bn, s, p, (1, 1, 1), 1, L.ACT_RELU, rr, name)
        plan.finalize()
        st = torch.cuda.current_stream().cuda_stream
        plan.run(st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn = plan.ops[-1][1]
        e0.record(); fn(st); e1.record(); torch.cuda.synchronize()
        m = plan.meta[-1]
        ms = e0.elapsed_time(e1)
        print("%-32s %8.1f us  %7.1f TFLOP/s  %7.1f GB/s" % (name, ms * 1e3, m["flops"] / ms / 1e9, m["bytes"] / ms / 1e6), flush=True)

main()
