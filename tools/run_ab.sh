#!/bin/bash
# A/B harness for env-switchable kernel variants: every variant is one fresh process (the switches are read once).
#   tools/run_ab.sh "<name>:<ENV=1 ENV2=2>" ...       results: gpurun_out/ab_<name>.json (+ per-launch times k_<name>.json)
mkdir -p gpurun_out
WL=${WORKLOAD:-slowfast_r50}
for spec in "$@"; do
  name="${spec%%:*}"; envs="${spec#*:}"
  [ "$envs" = "$spec" ] && envs=""
  echo "=== $name [$envs]"
  env $envs timeout 300 python bench.py --workload $WL --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline \
      --dump-kernels gpurun_out/k_${WL}_$name.json > gpurun_out/ab_${WL}_$name.json 2> gpurun_out/ab_${WL}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${WL}_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step %.4f"%d["ms_per_step"], "clips/s %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], "frac %.4f"%d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/ab_${WL}_$name.err").read()[-1500:])
PY
done
