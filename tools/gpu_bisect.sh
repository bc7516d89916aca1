#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_bisect.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "layernorm" 2>&1 | grep -E "passed|failed|timeout|Error|assert" | head -8 | tee -a $LOG
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "mvit" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 600 python bench.py --workload mvit_base_16x4 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mvit value %.1f ms/step %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['whole_model']['kernel_ms_by_kind'])" | tee -a $LOG
