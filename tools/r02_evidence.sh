#!/bin/bash
# Round-2 evidence run (one B200): per-workload bench lines (with cpu_baseline), the stock-PyTorch context arm, the
# transform bench, the ncu launch list + DRAM traffic of the bench step and ncu --set full captures of every hot kernel.
# Everything lands in gpurun_out/r02_*; tools/summarize_r02.py turns it into the tracked files under profiles/.
#   tools/r02_evidence.sh [bench|ncu|all]
what=${1:-all}
mkdir -p gpurun_out
if [ "$what" = bench ] || [ "$what" = all ]; then
  for w in slowfast_r50 mvit_base_16x4 x3d_m x3d_xs slow_r50 csn_r101 r2plus1d_r50; do
    timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --dump-kernels gpurun_out/r02_kernels_$w.json \
        > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
    tail -c 400 gpurun_out/r02_bench_$w.json | head -c 300; echo
  done
  for w in slowfast_r50 mvit_base_16x4 x3d_m; do
    timeout 300 python bench.py --impl torch-gpu --workload $w --steps 10 --warmup 3 > gpurun_out/r02_torchgpu_$w.json 2> gpurun_out/r02_torchgpu_$w.err
    cat gpurun_out/r02_torchgpu_$w.json | head -c 400; echo
  done
  timeout 400 python tools/bench_transform.py > gpurun_out/r02_transform.json 2> gpurun_out/r02_transform.err
  cat gpurun_out/r02_transform.json | head -c 1500; echo
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
  # launch list + DRAM bytes per launch of two bench steps (cold-cache, serialised: shares and bytes, not times)
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 260 --csv \
      --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
  # --set full captures: the dominant conv kernel on representative layers, the fused block, the stem kernel
  timeout 600 ncu --set full --clock-control none -k regex:"conv3d_igemm_kernel|conv3d_stem_rows|bottleneck_fused" \
      -o gpurun_out/r02_prof python tools/profile_layers.py res4_conv_a res2_conv_c res4_conv_b slow_stem fast_stem > gpurun_out/r02_prof.log 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:"bottleneck_fused" -c 2 -o gpurun_out/r02_prof_fused python tools/profile_fused.py res2 > gpurun_out/r02_prof_fused.log 2>&1
  # attention / depthwise / transform / (2+1)D at bench size
  timeout 400 ncu --set full --clock-control none -k regex:"attention" -s 4 -c 3 -o gpurun_out/r02_prof_attn python bench.py --workload mvit_base_16x4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_prof_attn.log 2>&1
  timeout 400 ncu --set full --clock-control none -k regex:"dwconv3d_tile" -s 30 -c 4 -o gpurun_out/r02_prof_dw python bench.py --workload x3d_m --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_prof_dw.log 2>&1
  timeout 400 ncu --set full --clock-control none -k regex:"clip_transform" -s 3 -c 2 -o gpurun_out/r02_prof_tr python tools/bench_transform.py > gpurun_out/r02_prof_tr.log 2>&1
  timeout 400 ncu --set full --clock-control none -k regex:"conv3d_igemm_kernel" -s 8 -c 3 -o gpurun_out/r02_prof_r2p1 python bench.py --workload r2plus1d_r50 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_prof_r2p1.log 2>&1
  # the reports are too big to travel back (64 MiB limit): export the raw page of each on the box, keep the small ones
  for r in gpurun_out/r02_prof*.ncu-rep; do
    ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
    [ $(stat -c %s $r) -gt 8000000 ] && rm -f $r
  done
  ls -la gpurun_out/r02_prof*
fi
