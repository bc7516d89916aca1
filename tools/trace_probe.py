"""Per-role wait accounting of the gather kernel (PVB200_TRACE) on the fast-pathway layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PVB200_TRACE"] = "1"
os.environ["PVB200_TRACE_BRIEF"] = "1"
import torch
import torch.nn as nn
from pytorchvideo_b200 import _lib as L
from pytorchvideo_b200.engine.plan import Plan

LAYERS = [
    ("fast_res2_conv_a_3x1x1_32to8", (8, 32, 32, 56, 56), 8, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),
    ("fast_res2_conv_b_1x3x3_8", (8, 8, 32, 56, 56), 8, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("fast_res2_conv_c_8to32_res", (8, 8, 32, 56, 56), 32, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
]
dev = torch.device("cuda:0")
for name, xs, co, k, s, p, use_res in LAYERS:
    plan = Plan(dev, L.PV_F16, True)
    x = torch.randn(xs, device=dev)
    xr = plan.emit_input_ncdhw(x, xs[1], xs[1])
    w = torch.randn(co, xs[1], *k) * 0.05
    bn = nn.BatchNorm3d(co).eval()
    rr = None
    if use_res:
        r = torch.randn(xs[0], co, xs[2], xs[3], xs[4], device=dev)
        rr = plan.emit_input_ncdhw(r, co, co)
    plan.emit_conv(xr, w, None, bn, s, p, (1, 1, 1), 1, L.ACT_RELU, rr, name)
    plan.finalize()
    st = torch.cuda.current_stream().cuda_stream
    os.environ.pop("PVB200_TRACE", None)
    plan.run(st); torch.cuda.synchronize()
    for dbg in sys.argv[1:] or ["64", "0"]:
        os.environ["PVB200_DEBUG"] = dbg
        os.environ["PVB200_TRACE"] = "1"
        sys.stderr.write("---- %s dbg=%s\n" % (name, dbg)); sys.stderr.flush()
        plan.ops[-1][1](st); torch.cuda.synchronize()
        plan.ops[-1][1](st); torch.cuda.synchronize()
        os.environ.pop("PVB200_TRACE", None)
