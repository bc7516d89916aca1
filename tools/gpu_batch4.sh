#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_batch4.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 900 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "conv3d and tcgen05"
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "f16" -s

run python tools/narrow_probe.py 0 64 128 160 164 36
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_slowfast_r50.json 2> gpurun_out/bench_slowfast_r50.err
grep -E "^(===|---)|passed|failed|rror|f16:|tiles/SM|clk/box" $LOG | tail -70
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_slowfast_r50.json').read().strip().splitlines()[-1])
print("slowfast value %.1f ms/step %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
