#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/gpu_mvit.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 900 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "layernorm or attention"
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "mvit" -s
grep -E "^(===|---)|passed|failed|rror|f16:|assert" $LOG | tail -40
