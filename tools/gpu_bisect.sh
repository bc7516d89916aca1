#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_bisect.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv3d" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "f16" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
timeout 300 python tools/narrow_probe.py 0 2>&1 | tee -a $LOG
for wl in slowfast_r50 x3d_m; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('narrowTMA $wl value %.1f ms/step %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['clocks'])" | tee -a $LOG
PVB200_GATHER_ALL=1 timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('gather $wl value %.1f ms/step %.3f e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a $LOG
done
