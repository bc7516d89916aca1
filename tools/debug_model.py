"""Run a workload's plan op by op with a synchronize after each launch; report the first failure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model_and_inputs, make_inputs
from pytorchvideo_b200.engine import compile_model

w = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
model, B, T, H, W, is_sf = build_model_and_inputs(w, B)
inp = make_inputs(B, T, H, W, is_sf, seed=1)
dev = torch.device("cuda:0")
din = [t.to(dev) for t in inp] if is_sf else inp.to(dev)
cm = compile_model(model, din, dtype="f16", use_graph=False)
for s, t in zip(cm.static_in, din if is_sf else [din]):
    s.copy_(t)
st = torch.cuda.current_stream().cuda_stream
for (name, fn), meta in zip(cm.plan.ops, cm.plan.meta):
    try:
        fn(st)
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED at op", name, meta["kind"], "->", str(e)[:300])
        sys.exit(1)
print("all", len(cm.plan.ops), "ops ok; out abs max", float(cm.output_view().abs().max()))
