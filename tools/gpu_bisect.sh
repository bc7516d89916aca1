#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_bisect.log
: > $LOG
echo "=== dbg 64 (TMA epilogue)" | tee -a $LOG
PVB200_DEBUG=64 timeout 120 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv3d and tcgen05" 2>&1 | grep -E "passed|failed|timeout|Error" | head -8 | tee -a $LOG
echo "=== sanitizer, direct epilogue" | tee -a $LOG
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv3d and tcgen05" 2>&1 | grep -v "^$" | head -60 | tee -a $LOG
