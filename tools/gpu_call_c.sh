#!/bin/bash
# ring2 (double-buffered staging for wide tiles without a residual): full GPU suite with it ON, layer sweep, whole-model A/B
mkdir -p gpurun_out
PVB200_EPI_RING2=1 timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/c_pytest_ring2.log 2>&1
echo "pytest(ring2) rc=$?"; grep -E "passed|failed" gpurun_out/c_pytest_ring2.log | tail -n 3; grep -E "^FAILED|^ERROR" gpurun_out/c_pytest_ring2.log | head
SWEEP_VARIANTS=base,ring2,ring2_no_store,ring2_no_math,ring2_ld64 timeout 300 python tools/epi_sweep.py nores conv_a conv_b conv_c mvit > gpurun_out/c_epi_sweep.jsonl 2> gpurun_out/c_epi_sweep.err
cat gpurun_out/c_epi_sweep.jsonl; tail -n 3 gpurun_out/c_epi_sweep.err
for w in slowfast_r50 mvit_base_16x4 x3d_m csn_r101 r2plus1d_r50; do
  for v in 0 1; do
    PVB200_EPI_RING2=$v timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --resident-only 2>> gpurun_out/c_bench.err | tail -n 1 | tee -a gpurun_out/c_bench.jsonl
  done
done
