"""A/B sweep of single SlowFast-R50 (B=8) Slow-pathway layers under the kernel's debug switches, timed as a CUDA graph of
REP identical launches (no host launch latency in the number).  One process: PVB200_DEBUG / PVB200_BN are read per launch.
   python tools/epi_sweep.py [layer-substring ...]
PVB200_DEBUG bits: 1 skip stores, 2 skip epilogue math, 4 producers skip loads, 32 MMA warp skips the MMAs,
512 single-buffer residual epilogue.  PVB200_BN = tile width override, PVB200_EPI_TEAMS = 1: one epilogue team (eight warps on one
tile) instead of two teams on alternate tiles for full-width residual tiles."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from pytorchvideo_b200 import _lib as L
from pytorchvideo_b200.engine.plan import Plan

LAYERS = [
    # name, (N,Ci,T,H,W), Co, k, s, p, residual
    ("res2_conv_a_256to64", (8, 256, 8, 56, 56), 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("res2_conv_b_1x3x3_64", (8, 64, 8, 56, 56), 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("res2_conv_c_64to256_res", (8, 64, 8, 56, 56), 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("res2_conv_c_64to256_nores", (8, 64, 8, 56, 56), 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("res3_conv_a_512to128", (8, 512, 8, 28, 28), 128, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("res3_conv_c_128to512_res", (8, 128, 8, 28, 28), 512, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("res4_conv_a_3x1x1_1024to256", (8, 1024, 8, 14, 14), 256, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),
    ("res4_conv_b_1x3x3_256", (8, 256, 8, 14, 14), 256, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("res4_conv_c_256to1024_res", (8, 256, 8, 14, 14), 1024, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
    ("res5_conv_a_3x1x1_2048to512", (8, 2048, 8, 7, 7), 512, (3, 1, 1), (1, 1, 1), (1, 0, 0), False),
    ("res5_conv_b_1x3x3_512", (8, 512, 8, 7, 7), 512, (1, 3, 3), (1, 1, 1), (0, 1, 1), False),
    ("res5_conv_c_512to2048_res", (8, 512, 8, 7, 7), 2048, (1, 1, 1), (1, 1, 1), (0, 0, 0), True),
]
ACTS = {"mvit_fc1_96to384_gelu": L.ACT_GELU, "mvit_fc1_384to1536_gelu": L.ACT_GELU, "mvit_qkv_96to288": L.ACT_NONE,
        "mvit_fc2_1536to384": L.ACT_NONE}
LAYERS += [
    ("mvit_fc1_96to384_gelu", (8, 96, 1, 1, 25089), 384, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("mvit_qkv_96to288", (8, 96, 1, 1, 25089), 288, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("mvit_fc1_384to1536_gelu", (8, 384, 1, 1, 1569), 1536, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
    ("mvit_fc2_1536to384", (8, 1536, 1, 1, 1569), 384, (1, 1, 1), (1, 1, 1), (0, 0, 0), False),
]
VARIANTS = [("base", {}), ("bn128", {"PVB200_BN": "128"}), ("bn64", {"PVB200_BN": "64"}),
            ("no_store", {"PVB200_DEBUG": "1"}), ("no_math", {"PVB200_DEBUG": "2"}),
            ("epi_only", {"PVB200_DEBUG": "36"}), ("epi_only_no_store", {"PVB200_DEBUG": "37"}),
            ("no_mma", {"PVB200_DEBUG": "32"}), ("single_buf", {"PVB200_DEBUG": "512"}),
            ("one_team", {"PVB200_EPI_TEAMS": "1"}), ("one_team_no_store", {"PVB200_EPI_TEAMS": "1", "PVB200_DEBUG": "1"}),
            ("one_team_no_math", {"PVB200_EPI_TEAMS": "1", "PVB200_DEBUG": "2"})]
if os.environ.get("SWEEP_VARIANTS"):
    VARIANTS = [v for v in VARIANTS if v[0] in os.environ["SWEEP_VARIANTS"].split(",")]
REP = 10


def main():
    only = sys.argv[1:]
    dev = torch.device("cuda:0")
    rows = []
    for name, xs, co, k, s, p, use_res in LAYERS:
        if only and not any(o in name for o in only):
            continue
        plan = Plan(dev, L.PV_F16, True)
        x = torch.randn(xs, device=dev)
        xr = plan.emit_input_ncdhw(x, xs[1], xs[1])
        w = torch.randn(co, xs[1], *k) * 0.05
        bn = nn.BatchNorm3d(co).eval()
        To = (xs[2] + 2 * p[0] - k[0]) // s[0] + 1
        Ho = (xs[3] + 2 * p[1] - k[1]) // s[1] + 1
        Wo = (xs[4] + 2 * p[2] - k[2]) // s[2] + 1
        rr = None
        if use_res:
            r = torch.randn(xs[0], co, To, Ho, Wo, device=dev)
            rr = plan.emit_input_ncdhw(r, co, co)
        plan.emit_conv(xr, w, None, bn, s, p, (1, 1, 1), 1, ACTS.get(name, L.ACT_RELU), rr, name)
        plan.finalize()
        st = torch.cuda.current_stream().cuda_stream
        plan.run(st)
        torch.cuda.synchronize()
        fn = plan.ops[-1][1]
        m = plan.meta[-1]
        out = {"layer": name, "gflop": m["flops"] / 1e9, "mb": m["bytes"] / 1e6}
        for vname, env in VARIANTS:
            if "PVB200_BN" in env and int(env["PVB200_BN"]) >= co:
                continue
            if (vname == "single_buf" or vname.startswith("one_team")) and not use_res:
                continue
            for kk in ("PVB200_DEBUG", "PVB200_BN", "PVB200_EPI_TEAMS"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            side = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, stream=side):
                    for _ in range(REP):
                        fn(torch.cuda.current_stream().cuda_stream)
            except Exception as e:      # a variant the kernel rejects
                out[vname] = None
                print(name, vname, "failed:", str(e)[:100], flush=True)
                continue
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / REP)
            out[vname] = round(best * 1e3, 2)
        for kk in ("PVB200_DEBUG", "PVB200_BN", "PVB200_EPI_TEAMS"):
            os.environ.pop(kk, None)
        print(json.dumps(out), flush=True)
        rows.append(out)
    return rows


if __name__ == "__main__":
    main()
