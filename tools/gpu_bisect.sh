#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_bisect.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "dwconv3d" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "f16 and x3d" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 600 python bench.py --workload x3d_m --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('x3d_m value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), d['whole_model']['kernel_ms_by_kind'])" | tee -a $LOG
for g in 1 2 3 4; do
PVB200_G=$g timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('G=$g slowfast value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), d['whole_model']['kernel_ms_by_kind'])" | tee -a $LOG
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('G=auto slowfast value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), d['whole_model']['kernel_ms_by_kind'])" | tee -a $LOG
