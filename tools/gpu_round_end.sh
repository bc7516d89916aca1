#!/bin/bash
# Round-end measurement pack: full GPU test-suite, transform bench, final bench lines (all workloads), ncu launch
# list of the bench command, ncu --set full captures of the dominant kernels.
mkdir -p gpurun_out
R=r01
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${R}_smoke.log
python tools/bench_transform.py > gpurun_out/${R}_transform.json 2> gpurun_out/${R}_transform.err; echo "transform exit $?"; cat gpurun_out/${R}_transform.json
python bench.py --steps 30 --warmup 5 --dump-kernels gpurun_out/${R}_kernels_slowfast.json > gpurun_out/${R}_bench_slowfast.json 2> gpurun_out/${R}_bench_slowfast.err; echo "bench exit $?"; cut -c1-1200 gpurun_out/${R}_bench_slowfast.json
for wl in x3d_m csn_r101 mvit_base_16x4 r2plus1d_r50 x3d_xs slow_r50; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --dump-kernels gpurun_out/${R}_kernels_$wl.json > gpurun_out/${R}_bench_$wl.json 2> gpurun_out/${R}_bench_$wl.err
  echo "bench $wl exit $?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 480 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:"conv3d_igemm" -o gpurun_out/${R}_prof python tools/profile_layers.py res4_conv_a res2_conv_c res4_conv_b fast_res2_conv_b fast_stem > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu --set full --clock-control none --import-source on -k regex:"dwconv3d_tile|attention_mma|clip_transform" -c 6 -o gpurun_out/${R}_prof_aux python bench.py --workload x3d_xs --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_aux.log 2>&1; echo "ncu aux exit $?"
python - <<'PY'
import json
for wl in ["slowfast", "x3d_m", "csn_r101", "mvit_base_16x4", "r2plus1d_r50", "x3d_xs", "slow_r50"]:
    try:
        d = json.loads(open('gpurun_out/r01_bench_%s.json' % wl).read().strip().splitlines()[-1])
        print(wl, "value %.1f ms/step %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["whole_model"]["kernel_ms_by_kind"], "roof %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(wl, "parse failed", e)
PY
ls -la gpurun_out | tail -5
