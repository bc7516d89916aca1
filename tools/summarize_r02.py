"""Turn the round-2 evidence run (tools/r02_evidence.sh, files gpurun_out/r02_*) into the tracked summaries under profiles/:
  r02_bench_<workload>.json, r02_kernels_<workload>.json, r02_torchgpu_<w>.json, r02_transform.json   (copies)
  r02_launch_list_summary.md, r02_launches.csv   per-kernel launch counts, time shares and DRAM bytes of two bench steps
  r02_traffic.json                                DRAM bytes per launch per kernel (bench.py reads it for roofline.traffic)
  r02_ncu_full_summary.md                         ncu --set full captures: duration, DRAM bytes, tensor pipe, L2 / DRAM throughput
  r02_sass_histogram.md                           tools/sass_histogram.py
    python tools/summarize_r02.py"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def short(name):
    n = name.split("(")[0].strip()
    if n.startswith("void "):
        n = n[5:]
    n = n.split("<")[0]
    return n.replace("pv::", "")


def copy_jsons():
    for f in glob.glob(os.path.join(G, "r02_bench_*.json")) + glob.glob(os.path.join(G, "r02_kernels_*.json")) + \
            glob.glob(os.path.join(G, "r02_torchgpu_*.json")) + [os.path.join(G, "r02_transform.json")]:
        if os.path.exists(f) and os.path.getsize(f) > 2:
            txt = open(f).read().strip().splitlines()[-1]
            try:
                json.loads(txt)
            except Exception:
                continue
            open(os.path.join(P, os.path.basename(f)), "w").write(txt + "\n")


def launch_list():
    src = os.path.join(G, "r02_launches.csv")
    if not os.path.exists(src):
        return
    per = {}
    for r in csv.reader(open(src)):
        if len(r) < 15 or not r[0].isdigit():
            continue
        d = per.setdefault(int(r[0]), {"kernel": short(r[4]), "grid": r[8]})
        try:
            d[r[12]] = float(r[14].replace(",", ""))
        except ValueError:
            pass
        d[r[12] + ".unit"] = r[13]
    rows = [per[k] for k in sorted(per)]

    def us(d):
        v, u = d.get("gpu__time_duration.sum", 0.0), d.get("gpu__time_duration.sum.unit", "ns")
        return v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)

    def mb(d, key):
        v, u = d.get(key, 0.0), d.get(key + ".unit", "byte")
        mul = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        return v * mul
    agg = {}
    for d in rows:
        a = agg.setdefault(d["kernel"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += us(d)
        a[2] += mb(d, "dram__bytes_read.sum")
        a[3] += mb(d, "dram__bytes_write.sum")
    tot = sum(a[1] for a in agg.values())
    out = ["# ncu launch list, round 2", "",
           "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 260 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`",
           "", "%d consecutive launches (~2.5 forward passes of SlowFast-8x8-R50, batch 8); per-launch times under ncu are cold-cache and serialised -" % len(rows),
           "compare SHARES, not absolutes.  DRAM bytes are per launch (mean).", "",
           "| kernel | launches | total us | share | DRAM read MB / launch | DRAM write MB / launch |", "|---|---|---|---|---|---|"]
    traffic = {}
    for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.1f %% | %.2f | %.2f |" % (k, n, t, 100.0 * t / tot, rd / n, wr / n))
        traffic[k] = (rd + wr) / n * 1e6
    open(os.path.join(P, "r02_launch_list_summary.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
    shutil.copy(src, os.path.join(P, "r02_launches.csv"))


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed.avg.per_cycle_active", "IPC"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def full_reports():
    out = ["# ncu `--set full --clock-control none` captures, round 2 (B200 sm_100a)", "",
           "Commands: tools/r02_evidence.sh (ncu section).  Times under ncu are cold-cache and serialised.", ""]
    # the box exports each capture's raw page (r02_prof*.raw.csv) because the .ncu-rep files exceed the copy-back limit
    reps = sorted(set(glob.glob(os.path.join(G, "r02_prof*.raw.csv")) + glob.glob(os.path.join(G, "r02_prof*.ncu-rep"))))
    seen = set()
    for rep in reps:
        stem = os.path.basename(rep).replace(".raw.csv", "").replace(".ncu-rep", "")
        if stem in seen:
            continue
        seen.add(stem)
        try:
            if rep.endswith(".csv"):
                raw = open(rep).read()
            else:
                raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        except Exception as e:
            out.append("%s: unreadable (%s)" % (os.path.basename(rep), e))
            continue
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        out += ["## %s" % os.path.basename(rep), "", "| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---|" * len(WANT)]
        for r in rows[2:]:
            cells = []
            for key, _ in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    cells.append("%s %s" % (r[i], units[i]) if units[i] not in ("", "%") else r[i])
                else:
                    cells.append("-")
            out.append("| %s | %s |" % (short(r[hdr.index("Kernel Name")]), " | ".join(cells)))
        out.append("")
    open(os.path.join(P, "r02_ncu_full_summary.md"), "w").write("\n".join(out) + "\n")


def sass():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_histogram.py")], capture_output=True, text=True)
    if res.returncode == 0:
        open(os.path.join(P, "r02_sass_histogram.md"), "w").write(res.stdout)


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    copy_jsons()
    launch_list()
    full_reports()
    sass()
    print("profiles/ updated:", sorted(f for f in os.listdir(P) if f.startswith("r02_")))
