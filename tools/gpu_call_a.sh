#!/bin/bash
# validation of the MViT fp32 trunk / fused pooling launches + per-layer dumps + epilogue sweep (one 1-GPU call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q -s \
    -k "layer or layernorm or attention or mvit or add_layernorm" > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "PARITY|passed|failed|Error|error" gpurun_out/a_pytest.log | tail -40
timeout 300 python bench.py --workload mvit_base_16x4 --steps 20 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/a_kernels_mvit.json > gpurun_out/a_bench_mvit.json 2> gpurun_out/a_bench_mvit.err
tail -c 1200 gpurun_out/a_bench_mvit.json | head -c 500; echo
PVB200_TRUNK32=0 timeout 300 python bench.py --workload mvit_base_16x4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_mvit_trunk16.json 2> gpurun_out/a_bench_mvit_trunk16.err
head -c 300 gpurun_out/a_bench_mvit_trunk16.json; echo
timeout 300 python bench.py --workload slowfast_r50 --steps 20 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/a_kernels_slowfast.json > gpurun_out/a_bench_slowfast.json 2> gpurun_out/a_bench_slowfast.err
head -c 300 gpurun_out/a_bench_slowfast.json; echo
timeout 200 python tools/epi_sweep.py > gpurun_out/a_epi_sweep.jsonl 2> gpurun_out/a_epi_sweep.err
cat gpurun_out/a_epi_sweep.jsonl; tail -3 gpurun_out/a_epi_sweep.err
