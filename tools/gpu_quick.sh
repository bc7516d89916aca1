#!/bin/bash
# quick GPU iteration: tcgen05 conv tests (isolated), f16 models, bench
mkdir -p gpurun_out
LOG=gpurun_out/gpu_quick.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 600 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
for i in 0 1 2 3 4 5 6 7 8; do
  run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "test_conv3d_bn_act and ${i}-f16-tcgen05"
done
run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "pool"
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "not f32" -s
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/kernels_slowfast.json > gpurun_out/bench_slowfast.json 2> gpurun_out/bench_slowfast.err
echo "bench exit $?" | tee -a $LOG
grep -E "^(===|---)|passed|failed|rror|f16:" $LOG | tail -40
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_slowfast.json').read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["whole_model"])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_slowfast.err').read()[-2000:])
PY
