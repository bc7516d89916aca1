#!/bin/bash
# Round-2 scaling curve on ONE 8-GPU B200 box (run as: gpurun --gpus 8 -- tools/r02_scale.sh).
# One process per GPU (torchrun, NCCL), bench.py's own barrier + max-over-ranks timing; the driver computes efficiency
# itself, these files are the builder's own curve: gpurun_out/r02_scale_{1,2,4,8}.json (SlowFast-8x8-R50 B=8/GPU, resident
# + e2e), r02_scale_x3d_m_{1,8}.json (X3D-M B=32/GPU) and the NCCL INFO lines of the 8-rank run.
mkdir -p gpurun_out
port=29511
run() {  # n workload out extra-env
  local n=$1 w=$2 out=$3
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/$out.json 2> gpurun_out/$out.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --workload $w --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/$out.json 2> gpurun_out/$out.err
    port=$((port + 1))
  fi
  tail -n 1 gpurun_out/$out.json | head -c 700; echo
}
nvidia-smi -L | head -8
for n in 1 2 4; do run $n slowfast_r50 r02_scale_$n; done
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH run 8 slowfast_r50 r02_scale_8
grep -E "NCCL INFO (Channel|Connected|comm .* rank|NVLS|Using network|Trees|Rings)" gpurun_out/r02_scale_8.err | head -60 > gpurun_out/r02_scale_8_nccl.txt
sed -i '/NCCL INFO/d' gpurun_out/r02_scale_8.err
run 1 x3d_m r02_scale_x3d_m_1
run 8 x3d_m r02_scale_x3d_m_8
