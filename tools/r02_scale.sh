#!/bin/bash
# Round-2 scaling curve on ONE multi-GPU B200 box (run as: gpurun --gpus 8 -- tools/r02_scale.sh 1 8 4 2).
# One process per GPU (torchrun, NCCL), bench.py's own barrier + max-over-ranks timing; the driver computes efficiency
# itself, these files are the builder's own curve: gpurun_out/r02_scale_{N}.json (SlowFast-8x8-R50 B=8/GPU, resident
# + e2e) and the NCCL INFO lines of the largest run.  Ns after the first two are skipped once BUDGET_S seconds are spent
# (an 8-GPU box is charged 8x).
mkdir -p gpurun_out
port=29511
t0=$(date +%s)
BUDGET_S=${BUDGET_S:-150}
run() {  # n workload out
  local n=$1 w=$2 out=$3
  if [ "$n" = 1 ]; then
    timeout 200 python bench.py --gpus 1 --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/$out.json 2> gpurun_out/$out.err
  else
    NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
        --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
        > gpurun_out/$out.json 2> gpurun_out/$out.err
    port=$((port + 1))
    grep -E "NCCL INFO (Channel|Connected|comm .* rank|NVLS|Using network|Trees|Rings)" gpurun_out/$out.err | head -40 > gpurun_out/${out}_nccl.txt
    sed -i '/NCCL INFO/d' gpurun_out/$out.err
  fi
  echo "N=$n rc=$? $(($(date +%s) - t0))s"; tail -n 1 gpurun_out/$out.json | head -c 400; echo; tail -n 3 gpurun_out/$out.err
}
nvidia-smi -L | head -8
i=0
for n in "$@"; do
  i=$((i + 1))
  if [ $i -gt 2 ] && [ $(($(date +%s) - t0)) -gt $BUDGET_S ]; then echo "skip N=$n (budget)"; continue; fi
  run $n slowfast_r50 r02_scale_$n
done
