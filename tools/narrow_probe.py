"""Isolate what bounds the narrow (fast-pathway) conv layers: time each layer as a CUDA graph of REPS
launches (so the host is out of the picture) under the PVB200_DEBUG skip masks
(1 = no TMA stores, 2 = no epilogue math, 4 = producers issue no loads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from pytorchvideo_b200 import _lib as L
from pytorchvideo_b200.engine.plan import Plan

LAYERS = [
    ("fast_res2_conv_a_3x1x1_32to8", (8, 32, 32, 56, 56), 8, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),
    ("fast_res2_conv_b_1x3x3_8", (8, 8, 32, 56, 56), 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("fast_res2_conv_c_8to32_res", (8, 8, 32, 56, 56), 32, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("fast_res3_conv_b_1x3x3_16", (8, 16, 32, 28, 28), 16, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("slow_res2_conv_c_64to256_res", (8, 64, 8, 56, 56), 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("slow_res2_conv_a_256to64", (8, 256, 8, 56, 56), 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("fast_stem_5x7x7", (8, 3, 32, 224, 224), 8, (5, 7, 7), (1, 2, 2), (2, 3, 3), False),
]
REPS = 10


def main():
    masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 5, 7]
    dev = torch.device("cuda:0")
    for name, xs, co, k, s, p, use_res in LAYERS:
        plan = Plan(dev, L.PV_F16, True)
        x = torch.randn(xs, device=dev)
        xr = plan.emit_input_ncdhw(x, xs[1], 4 if xs[1] <= 4 else xs[1])
        w = torch.randn(co, xs[1], *k) * 0.05
        bn = nn.BatchNorm3d(co).eval()
        To = (xs[2] + 2 * p[0] - k[0]) // s[0] + 1; Ho = (xs[3] + 2 * p[1] - k[1]) // s[1] + 1; Wo = (xs[4] + 2 * p[2] - k[2]) // s[2] + 1
        rr = None
        if use_res:
            r = torch.randn(xs[0], co, To, Ho, Wo, device=dev)
            rr = plan.emit_input_ncdhw(r, co, co)
        plan.emit_conv(xr, w, None, bn, s, p, (1, 1, 1), 1, L.ACT_RELU, rr, name)
        plan.finalize()
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            plan.run(stream.cuda_stream)
        torch.cuda.synchronize()
        # the conv proper = the ops emitted by emit_conv (last 1 or 2: taps + tapsum)
        n_conv = 2 if name.startswith("fast_stem") else 1
        fns = [op[1] for op in plan.ops[-n_conv:]][:1]     # time the GEMM kernel only
        tiles = (xs[0] * To * Ho * Wo + 127) // 128
        out = []
        for m in masks:
            os.environ["PVB200_DEBUG"] = str(m)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for _ in range(REPS):
                    for fn in fns:
                        fn(torch.cuda.current_stream().cuda_stream)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / REPS
            out.append("dbg%d %7.1f us" % (m, us))
        os.environ["PVB200_DEBUG"] = "0"
        print("%-30s tiles/SM %5.1f | %s" % (name, tiles / 148.0, " | ".join(out)), flush=True)


main()
