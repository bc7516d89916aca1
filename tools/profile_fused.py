"""Run the fused Fast-pathway blocks of SlowFast-R50 (B=8) once each, for ncu:
    ncu --set full --import-source on -k regex:bottleneck_fused -o gpurun_out/r02_fused python tools/profile_fused.py [res2|res3|res4]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchvideo_b200 import testing as TS
from pytorchvideo_b200.engine import compile_model
from pytorchvideo_b200.models.resnet import create_bottleneck_block, create_res_block

CFG = {"res2": (8, 32, 56, 56, 32, 8, 32, 1), "res2_0": (8, 32, 56, 56, 8, 8, 32, 1), "res3": (8, 32, 28, 28, 64, 16, 64, 1),
       "res3_0": (8, 32, 56, 56, 32, 16, 64, 2), "res4": (8, 32, 14, 14, 128, 32, 128, 1)}


def main():
    for name in (sys.argv[1:] or ["res2", "res3", "res4"]):
        N, T, H, W, cin, cmid, cout, s = CFG[name]
        blk = create_res_block(dim_in=cin, dim_inner=cmid, dim_out=cout, bottleneck=create_bottleneck_block,
                               conv_a_kernel_size=(3, 1, 1), conv_a_stride=(1, 1, 1), conv_a_padding=(1, 0, 0), conv_b_stride=(1, s, s))
        blk = TS.randomize_model(blk, seed=1).eval()
        x = torch.randn(N, cin, T, H, W).cuda()
        cm = compile_model(blk, x, dtype="f16", use_graph=False)
        cm(x); torch.cuda.synchronize()
        i = [k for k, m in enumerate(cm.plan.meta) if m["kind"] == "fused_block"][0]
        fn = cm.plan.ops[i][1]
        st = torch.cuda.current_stream().cuda_stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(st); e1.record(); torch.cuda.synchronize()
        m = cm.plan.meta[i]
        ms = e0.elapsed_time(e1)
        print("%-8s %8.1f us  %6.1f GB/s  %5.1f TFLOP/s" % (name, ms * 1e3, m["bytes"] / ms / 1e6, m["flops"] / ms / 1e9), flush=True)


main()
