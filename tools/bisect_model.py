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


def main():
    case = sys.argv[1]
    dtype = sys.argv[2] if len(sys.argv) > 2 else "f16"
    model, inp, is_sf = TS.build_case(case, PH)
    nb = len(model.blocks)
    x_cpu = inp
    for k in range(1, nb):          # the head needs the full net; prefixes end with a feature map
        sub = copy.copy(model)
        sub.__dict__ = dict(model.__dict__)
        sub._modules = dict(model._modules)
        sub._modules["blocks"] = nn.ModuleList(list(model.blocks)[:k])
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
        print("blocks[:%d] %-28s out %s  max|d|/max|ref| = %.3e" % (k, type(model.blocks[k - 1]).__name__, tuple(ref.shape), err), flush=True)


main()
