#!/bin/bash
# Round-2 final evidence refresh on one B200 (the subset of tools/r02_evidence.sh that changed since the last full run):
# full GPU test log, bench lines (with cpu_baseline) of the three headline workloads, the ncu launch list + DRAM bytes of the SlowFast
# bench step, and an ncu --set full capture of the new add_layernorm kernel.  tools/summarize_r02.py turns gpurun_out/r02_* into profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed" gpurun_out/r02_pytest_gpu.log | tail -n 2; grep -E "^FAILED|^ERROR" gpurun_out/r02_pytest_gpu.log | head
for w in slowfast_r50 mvit_base_16x4 x3d_m; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --dump-kernels gpurun_out/r02_kernels_$w.json \
      > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
  tail -c 600 gpurun_out/r02_bench_$w.json | head -c 300; echo
done
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 260 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"add_layernorm|layernorm_reg" -s 6 -c 2 -o gpurun_out/r02_prof_ln \
    python bench.py --workload mvit_base_16x4 --steps 1 --warmup 3 --resident-only > gpurun_out/r02_prof_ln.log 2>&1
ncu -i gpurun_out/r02_prof_ln.ncu-rep --page raw --csv > gpurun_out/r02_prof_ln.raw.csv 2>/dev/null
[ -f gpurun_out/r02_prof_ln.ncu-rep ] && [ $(stat -c %s gpurun_out/r02_prof_ln.ncu-rep) -gt 8000000 ] && rm -f gpurun_out/r02_prof_ln.ncu-rep
SWEEP_VARIANTS=base,one_team,one_team_no_store,one_team_no_math,epi_only timeout 200 python tools/epi_sweep.py _res > gpurun_out/r02_epi_sweep_teams.jsonl 2> gpurun_out/r02_epi_sweep_teams.err
cat gpurun_out/r02_epi_sweep_teams.jsonl; tail -n 2 gpurun_out/r02_epi_sweep_teams.err
ls -la gpurun_out | head -30
