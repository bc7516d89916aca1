#!/bin/bash
# full validation + all benches after the chunked TMA pipeline / elected-lane issue / direct epilogue
mkdir -p gpurun_out
LOG=gpurun_out/gpu_batch5.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 1200 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
run python -m pytest tests -q -m gpu -x
for wl in slowfast_r50 x3d_m csn_r101 mvit_base_16x4 x3d_xs r2plus1d_r50 slow_r50; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/kernels_$wl.json > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl exit $?" | tee -a $LOG
done
run python tools/narrow_probe.py 0
grep -E "^(===|---)|passed|failed|rror|tiles/SM|bench " $LOG | tail -40
python - <<'PY'
import json
for wl in ["slowfast_r50", "x3d_m", "csn_r101", "mvit_base_16x4", "x3d_xs", "r2plus1d_r50", "slow_r50"]:
    try:
        d = json.loads(open('gpurun_out/bench_%s.json' % wl).read().strip().splitlines()[-1])
        print(wl, "value %.1f ms/step %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["whole_model"]["kernel_ms_by_kind"], "roof %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(wl, "bench parse failed", e); print(open('gpurun_out/bench_%s.err' % wl).read()[-1500:])
PY
