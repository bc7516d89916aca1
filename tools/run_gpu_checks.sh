#!/bin/bash
# Runs the GPU test-suite in isolated processes (a CUDA trap poisons its process only) and
# keeps full logs under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/gpu_checks.log
: > $LOG
run() { echo "=== $*" | tee -a $LOG; timeout 600 "$@" >> $LOG 2>&1; echo "--- exit $?" | tee -a $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run python __graft_entry__.py
run python -m pytest tests/test_gpu_transforms.py -q -m gpu
run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "direct or pool or layernorm or attention"
for i in 0 1 2 3 4 5 6 7 8; do
  run python -m pytest tests/test_gpu_ops.py -q -m gpu -k "test_conv3d_bn_act and ${i}-f16-tcgen05"
done
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "f32" -s
run python -m pytest tests/test_gpu_models.py -q -m gpu -k "not f32" -s
run python __graft_entry__.py smoke
grep -E "^(===|---)|passed|failed|error" $LOG | tail -60
