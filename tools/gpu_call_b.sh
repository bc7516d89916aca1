#!/bin/bash
# LSU epilogue experiment: correctness of the op tests with PVB200_EPI_LSU=1, layer sweep, whole-model A/B
mkdir -p gpurun_out
PVB200_EPI_LSU=1 timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv3d_bn_act or homogeneous" > gpurun_out/b_pytest_lsu.log 2>&1
echo "pytest(lsu) rc=$?"; tail -n 4 gpurun_out/b_pytest_lsu.log
SWEEP_VARIANTS=base,lsu,lsu_ld64,lsu_no_store,lsu_no_math,lsu_epi_only,lsu_bn128 timeout 300 python tools/epi_sweep.py > gpurun_out/b_epi_sweep.jsonl 2> gpurun_out/b_epi_sweep.err
cat gpurun_out/b_epi_sweep.jsonl; tail -n 3 gpurun_out/b_epi_sweep.err
for w in slowfast_r50 mvit_base_16x4 x3d_m; do
  for lsu in 0 1; do
    PVB200_EPI_LSU=$lsu timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --resident-only 2>> gpurun_out/b_bench.err | tail -n 1 | tee -a gpurun_out/b_bench.jsonl
  done
done
PVB200_EPI_LSU=1 PVB200_DEBUG=4096 timeout 200 python bench.py --workload slowfast_r50 --steps 20 --warmup 5 --resident-only 2>> gpurun_out/b_bench.err | tail -n 1 | tee -a gpurun_out/b_bench.jsonl
PVB200_EPI_LSU=1 timeout 400 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "f16_tensor_core and (c2_slowfast or mvit_base_8x112_f16w or c1_x3d or slow_r50_f16w or r2plus1d)" > gpurun_out/b_pytest_models_lsu.log 2>&1
echo "pytest(models, lsu) rc=$?"; tail -n 4 gpurun_out/b_pytest_models_lsu.log
