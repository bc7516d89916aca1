#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_batch2.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 900 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention or 9- or 10- or 11-"
run python -m pytest tests/test_gpu_transforms.py -q -m gpu
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "f16 and (x3d or csn or mvit)" -s
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "f32 and (x3d_xs or mvit)" -s
grep -E "^(===|---)|passed|failed|rror|f16:|assert" $LOG | tail -30
for w in x3d_m mvit_base_16x4 csn_r101; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels_$w.json > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
    print("$w", "value %.1f clips/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["whole_model"]["kernel_ms_by_kind"])
except Exception as e:
    print("$w parse fail", e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
python tools/bench_transform.py 2>&1 | tail -1 | cut -c1-700
