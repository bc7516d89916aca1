#!/bin/bash
# Round-end measurement pack: transform bench, final bench line, ncu launch list, ncu full captures.
mkdir -p gpurun_out
python tools/bench_transform.py > gpurun_out/r01_transform.json 2> gpurun_out/r01_transform.err; echo "transform exit $?"; cat gpurun_out/r01_transform.json
python bench.py --steps 30 --warmup 5 --dump-kernels gpurun_out/r01_kernels_slowfast.json > gpurun_out/r01_bench_slowfast.json 2> gpurun_out/r01_bench_slowfast.err; echo "bench exit $?"; cut -c1-900 gpurun_out/r01_bench_slowfast.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_ncu_bench.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:"conv3d_igemm|clip_transform" -o gpurun_out/r01_prof python tools/profile_layers.py res4_conv_a res2_conv_c res4_conv_b fast_res2_conv_b slow_stem > gpurun_out/r01_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | tail -12
