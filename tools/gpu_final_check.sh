#!/bin/bash
# what the driver does at round end: GPU tests, smoke, default bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 900 python bench.py 2>gpurun_out/final_bench.err | tail -1 > gpurun_out/final_bench.json; python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read())
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "cpu_baseline", "clocks", "gpu_launches")})
print(d["roofline"])
PY
