"""Turn the round-end ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_profiles.py r01

Inputs : gpurun_out/<R>_launches.csv (ncu --metrics gpu__time_duration.sum launch list of bench.py),
         gpurun_out/<R>_prof.ncu-rep, <R>_prof_aux.ncu-rep (ncu --set full captures), bench / kernel JSONs.
Outputs: profiles/<R>_launch_list_summary.md, profiles/<R>_ncu_full_summary.md and copies of the JSON/CSV files.
"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


try:
    HBM_GBS = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6572.5))
except Exception:
    HBM_GBS = 6572.5


def short(name):
    n = name.split("(")[0].strip()
    for pre in ("void ",):
        if n.startswith(pre):
            n = n[len(pre):]
    return n.split("<")[0]


def launch_list():
    src = os.path.join(G, "%s_launches.csv" % R)
    rows = []
    for r in csv.reader(open(src)):
        if len(r) > 14 and r[0].isdigit():
            rows.append((int(r[0]), short(r[4]), r[8], float(r[14]) / 1000.0))
    tot = sum(r[3] for r in rows)
    agg = {}
    for _, k, _, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    out = ["# ncu launch list, %s: `ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`" % R,
           "", "%d consecutive launches (~4 forward passes of SlowFast-8x8-R50, batch 8); cold-cache, serialised times - shares only." % len(rows),
           "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.1f %% |" % (k, n, us, 100.0 * us / tot))
    # one forward pass in launch order (between two layout-conversion launches of the slow pathway input)
    starts = [i for i, r in enumerate(rows) if "padw" in r[1]]
    if len(starts) >= 3:
        a, b = starts[0], starts[2]
        seg = rows[a:b]
        out += ["", "One forward pass in launch order (%d launches, %.1f us summed; bench.py measures %.0f us per CUDA-graph replay):" % (
            len(seg), sum(r[3] for r in seg), bench_ms() * 1e3), "", "```"]
        line = []
        for r in seg:
            line.append("%s:%.0f" % ({"pv::conv3d_igemm_kernel": "T", "pv::conv3d_igemm_gather_kernel": "G"}.get(r[1], r[1].replace("pv::", "")[:14]), r[3]))
            if len(line) == 12:
                out.append(" ".join(line)); line = []
        if line:
            out.append(" ".join(line))
        out += ["```", "(T = TMA-fed tcgen05 kernel, G = gather-fed tcgen05 kernel; numbers are microseconds)"]
    open(os.path.join(P, "%s_launch_list_summary.md" % R), "w").write("\n".join(out) + "\n")
    shutil.copy(src, os.path.join(P, "%s_launches.csv" % R))


def bench_ms():
    try:
        d = json.loads(open(os.path.join(G, "%s_bench_slowfast.json" % R)).read().strip().splitlines()[-1])
        return d["ms_per_step"]
    except Exception:
        return float("nan")


def raw_table(rep):
    res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    hdr = rows[0]

    def col(name, exact=True):
        for i, h in enumerate(hdr):
            if (h == name) if exact else (name in h):
                return i
        return None
    cols = {
        "name": col("Kernel Name"), "grid": col("launch__grid_size"), "regs": col("launch__registers_per_thread"),
        "dur": col("gpu__time_duration.sum"), "rd": col("dram__bytes_read.sum"), "wr": col("dram__bytes_write.sum"),
        "tensor": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "l2": col("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram": col("dram__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "sm": col("sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "smem": col("launch__shared_mem_per_block_dynamic", False),
    }
    units = rows[1]
    out = []
    for r in rows[2:]:
        def f(key, scale=1.0):
            i = cols[key]
            if i is None or r[i] == "":
                return float("nan")
            v = float(r[i].replace(",", ""))
            u = units[i]
            if key == "dur":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            if key in ("rd", "wr"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            return v * scale
        rec = {"name": short(r[cols["name"]]), "grid": r[cols["grid"]], "regs": r[cols["regs"]], "dur": f("dur"), "rd": f("rd"),
               "wr": f("wr"), "tensor": f("tensor"), "l2": f("l2"), "dram": f("dram"), "sm": f("sm"), "smem": f("smem")}
        if rec["dram"] != rec["dram"]:      # not collected on this driver: derive from bytes / duration against the measured copy peak
            rec["dram"] = (rec["rd"] + rec["wr"]) / rec["dur"] * 1e3 / HBM_GBS * 100.0
        out.append(rec)
    return out


LAYER_NOTES = {  # tools/profile_layers.py order for the captured selection; GFLOP = algorithmic
    "res4_conv_a": ("res4 conv_a 1024->256 (3,1,1) (TMA)", 19.73, 31.5),
    "res2_conv_c": ("res2 conv_c 64->256 pointwise + residual + ReLU (TMA)", 6.58, 231.2),
    "res4_conv_b": ("res4 conv_b 256->256 (1,3,3) (TMA)", 14.80, 14.0),
    "fast_res2_conv_b": ("fast res2 conv_b 8->8 (1,3,3) (gather-fed)", 0.92, 25.7),
    "fast_stem": ("fast stem 3->8 (5,7,7)/s(1,2,2): window-mode TMA taps GEMM (N = 5x8)", 37.76, 334.0),
}


def full_summary():
    rep = os.path.join(G, "%s_prof.ncu-rep" % R)
    if not os.path.exists(rep):
        return
    t = raw_table(rep)
    # profile_layers.py launches each conv twice (warm-up, then timed); LAYERS order is fixed
    order = ["fast_stem", "res2_conv_c", "res4_conv_a", "res4_conv_b", "fast_res2_conv_b"]
    out = ["# ncu `--set full --clock-control none --import-source on` captures, %s (B200 sm_100a)" % R, "",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:\"conv3d_igemm\" -o gpurun_out/%s_prof python tools/profile_layers.py res4_conv_a res2_conv_c res4_conv_b fast_res2_conv_b fast_stem`" % R,
           "(each layer = one SlowFast-8x8-R50 convolution at batch 8, launched twice; the second launch is listed).  Per-launch times under ncu are cold-cache and serialised - compare shares, not absolutes.",
           "", "| layer | kernel | duration (us) | DRAM read (MB) | DRAM write (MB) | algorithmic MB | algorithmic GFLOP | TFLOP/s | tensor pipe active % (of active cycles) | L2 thr % | DRAM thr % | grid | regs/thread |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    convs = [k for k in t if "igemm" in k["name"]]
    # every layer is launched twice (second one listed); the stem's first launch precedes the capture window
    picks = [convs[0]] + [convs[i + 1] for i in range(1, len(convs) - 1, 2)] if len(convs) % 2 else [convs[i + 1] for i in range(0, len(convs) - 1, 2)]
    for key, k in zip(order, picks):
        desc, gf, mb = LAYER_NOTES[key]
        out.append("| %s | %s | %.1f | %.1f | %.1f | %.1f | %.2f | %.0f | %.1f | %.1f | %.1f | %s | %s |" % (
            desc, k["name"].replace("pv::", ""), k["dur"], k["rd"], k["wr"], mb, gf, gf / k["dur"] * 1e3, k["tensor"], k["l2"], k["dram"], k["grid"], k["regs"]))
    aux = os.path.join(G, "%s_prof_aux.ncu-rep" % R)
    if os.path.exists(aux):
        ta = raw_table(aux)
        out += ["", "Auxiliary kernels (`ncu --set full -k regex:\"dwconv3d_tile|attention_mma|clip_transform\" -c 6 python bench.py --workload x3d_xs --steps 1 --warmup 1`, X3D-XS batch 32):", "",
                "| kernel | duration (us) | DRAM read (MB) | DRAM write (MB) | SM thr % | L2 thr % | DRAM thr % | grid | regs/thread |", "|---|---|---|---|---|---|---|---|---|"]
        for k in ta:
            out.append("| %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %s |" % (k["name"].replace("pv::", ""), k["dur"], k["rd"], k["wr"], k["sm"], k["l2"], k["dram"], k["grid"], k["regs"]))
    out += ["", "Reading (details and the probe measurements behind it in DESIGN.md section 9):",
            "* deep-K, wide-N layers (res4) keep the tensor pipe ~40-45 % active on 98 of 148 SMs (M = 12544 rows = 98 tiles); the execution counters of the same capture (`profiles/r01_source_counters.md`) show the producer never waiting for a free slot and the MMA warp hardly waiting for data: the single TMA-issuing warp paces these layers (~1460 clk per k-block vs 512 clk of MMAs) - round 2: split A/B issue over two warps, then cta_group::2 / multicast;",
            "* narrow-N layers are bound by the issue/handshake cost of the warp-specialised ring (~400-800 clk per barrier round, tools/probe/umma_issue.cu) - several k-blocks per pipeline stage took the fast stem from 559 to ~290 us, the gather-fed fast-pathway layers are still ~50 us each (a latency ring, ~1 us per 128-row tile);",
            "* res2 conv_c (+residual) is epilogue-bound (the MMA warp waits for a free accumulator 68x per tile): residual prefetch / double staging is the fix;",
            "* DRAM traffic is at or below the algorithmic bytes everywhere (activations are L2-resident across consecutive layers); no wasted re-reads."]
    open(os.path.join(P, "%s_ncu_full_summary.md" % R), "w").write("\n".join(out) + "\n")


def copies():
    for f in glob.glob(os.path.join(G, "%s_bench_*.json" % R)) + glob.glob(os.path.join(G, "%s_kernels_slowfast.json" % R)) + \
            [os.path.join(G, "%s_transform.json" % R), os.path.join(G, "%s_pytest_gpu.log" % R)]:
        if os.path.exists(f) and os.path.getsize(f) > 0:
            shutil.copy(f, P)


os.makedirs(P, exist_ok=True)
copies()
launch_list()
full_summary()
print(open(os.path.join(P, "%s_launch_list_summary.md" % R)).read()[:3000])
print(open(os.path.join(P, "%s_ncu_full_summary.md" % R)).read())
