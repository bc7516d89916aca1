#!/bin/bash
mkdir -p gpurun_out
for w in x3d_m mvit_base_16x4 x3d_xs csn_r101 r2plus1d_r50; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --dump-kernels gpurun_out/kernels_$w.json > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "$w exit $?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
    print("$w", "value %.1f clips/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], d["roofline"]["kernel"], "frac %.3f" % d["roofline"]["frac"], d["whole_model"]["kernel_ms_by_kind"])
except Exception as e:
    print("parse fail", e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
