"""Find the first block of a model whose CUDA output leaves the oracle: runs the prefixes blocks[:k] of a
MODEL_CASES entry through the engine and the CPU oracle and prints max|d|/max|ref| per prefix.
    python tools/bisect_model.py x3d_l [f16|f32]"""
import copy
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from oracle.interp import oracle_forward
from pytorchvideo_b200 import testing as TS
from pytorchvideo_b200.engine import compile_model
import pytorchvideo_b200.models.hub as PH


def _shallow(m):
    c = copy.copy(m)
    c.__dict__ = dict(m.__dict__)
    c._modules = dict(m._modules)
    return c


def main():
    """argv: case [f16|f32] [stage index: bisect the res_blocks of blocks[stage] instead of whole blocks]"""
    case = sys.argv[1]
    dtype = sys.argv[2] if len(sys.argv) > 2 else "f16"
    stage = int(sys.argv[3]) if len(sys.argv) > 3 else None
    model, inp, is_sf = TS.build_case(case, PH)
    nb = len(model.blocks)
    x_cpu = inp
    steps = range(1, nb) if stage is None else range(1, len(model.blocks[stage].res_blocks) + 1)
    for k in steps:          # the head needs the full net; prefixes end with a feature map
        sub = _shallow(model)
        if stage is None:
            sub._modules["blocks"] = nn.ModuleList(list(model.blocks)[:k])
        else:
            st = _shallow(model.blocks[stage])
            st._modules["res_blocks"] = nn.ModuleList(list(model.blocks[stage].res_blocks)[:k])
            sub._modules["blocks"] = nn.ModuleList(list(model.blocks)[:stage] + [st])
        ref = oracle_forward(sub, x_cpu)
        dev_in = [t.cuda() for t in inp] if is_sf else inp.cuda()
        try:
            cm = compile_model(sub, dev_in, dtype=dtype, use_graph=False)
            out = cm(dev_in)
            out = out.float().cpu()
        except NotImplementedError as e:
            print("blocks[:%d]: lowering stops here (%s)" % (k, e))
            break
        if isinstance(ref, list):
            print("blocks[:%d]: multi-pathway output, skipped" % k)
            continue
        err = float((out - ref).abs().max()) / max(float(ref.abs().max()), 1e-9)
        tag = "blocks[:%d]" % k if stage is None else "blocks[%d].res_blocks[:%d]" % (stage, k)
        print("%s out %s  max|d|/max|ref| = %.3e  mean|d|/max|ref| = %.3e" % (tag, tuple(ref.shape), err,
              float((out - ref).abs().mean()) / max(float(ref.abs().max()), 1e-9)), flush=True)


main()
