"""Debug helper: run ONE res block (its branch2 BottleneckBlock) of a model on the oracle's own input to that block and
compare with the oracle - with the squeeze-excitation on / replaced by Identity.
    python tools/debug_block.py x3d_l 4 6"""
import copy
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from oracle.interp import oracle_forward
from pytorchvideo_b200 import testing as TS
from pytorchvideo_b200.engine import compile_model
import pytorchvideo_b200.models.hub as PH


def shallow(m):
    c = copy.copy(m)
    c.__dict__ = dict(m.__dict__)
    c._modules = dict(m._modules)
    return c


def main():
    case, stage, blk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    model, inp, _ = TS.build_case(case, PH)
    pre = shallow(model)
    st = shallow(model.blocks[stage])
    st._modules["res_blocks"] = nn.ModuleList(list(model.blocks[stage].res_blocks)[:blk])
    pre._modules["blocks"] = nn.ModuleList(list(model.blocks)[:stage] + ([st] if blk > 0 else []))
    x = oracle_forward(pre, inp)
    print("input to block: shape %s  max %.3f  mean|x| %.3f" % (tuple(x.shape), float(x.abs().max()), float(x.abs().mean())))
    block = model.blocks[stage].res_blocks[blk]
    br = block.branch2
    for variant in ("full", "no_se", "res_block"):
        m = br
        if variant == "no_se" and isinstance(br.norm_b, nn.Sequential):
            m = shallow(br)
            nb = nn.Sequential(br.norm_b[0], nn.Identity())
            m._modules["norm_b"] = nb
        if variant == "res_block":
            m = block
        ref = oracle_forward(m, x)
        cm = compile_model(m, x.cuda(), dtype="f16", use_graph=False)
        out = cm(x.cuda()).float().cpu()
        d = (out - ref).abs()
        per_c = d.amax(dim=(0, 2, 3, 4)) / float(ref.abs().max())
        print("%-9s max|d|/max|ref| %.3e  mean %.3e  worst channels %s" % (
            variant, float(d.max()) / float(ref.abs().max()), float(d.mean()) / float(ref.abs().max()),
            [(int(i), "%.1e" % float(per_c[i])) for i in torch.argsort(per_c, descending=True)[:6]]))
        names = [n for n, _ in cm.plan.ops]
        print("          ops:", names)


main()
