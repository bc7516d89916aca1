#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_bisect.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv3d and tcgen05" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "f16" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 200 python tools/profile_layers.py res4 res5 res2_conv_b 2>&1 | tee -a $LOG
for wl in slowfast_r50 slow_r50 r2plus1d_r50; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']), d['whole_model']['kernel_ms_by_kind'])" | tee -a $LOG
done
