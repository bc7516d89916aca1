#!/bin/bash
# batch 3: TMA-tiled depthwise kernel + pipelined e2e + narrow-layer probe
mkdir -p gpurun_out
LOG=gpurun_out/gpu_batch3.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 900 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "dwconv3d or pool or (conv3d and direct)"
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "not f32" -s
for wl in slowfast_r50 x3d_m csn_r101 mvit_base_16x4 x3d_xs; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/kernels_$wl.json > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl exit $?" | tee -a $LOG
done
PVB200_DW_SIMT=1 timeout 600 python bench.py --workload x3d_m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x3d_m_simt.json 2> gpurun_out/bench_x3d_m_simt.err
run python tools/narrow_probe.py
grep -E "^(===|---)|passed|failed|rror|f16:|tiles/SM|bench " $LOG | tail -60
python - <<'PY'
import json
for wl in ["slowfast_r50", "x3d_m", "x3d_m_simt", "csn_r101", "mvit_base_16x4", "x3d_xs"]:
    try:
        d = json.loads(open('gpurun_out/bench_%s.json' % wl).read().strip().splitlines()[-1])
        print(wl, "value %.1f ms/step %.3f e2e %.1f (serial %.3f ms, piped %.3f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("serial_ms_per_step", 0), d["e2e"]["ms_per_step"]), d["whole_model"]["kernel_ms_by_kind"])
    except Exception as e:
        print(wl, "bench parse failed", e); print(open('gpurun_out/bench_%s.err' % wl).read()[-1500:])
PY
