timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -m gpu -x -k "conv3d or f16 or full_size" 2>&1 | tail -2
for wl in slowfast_r50 x3d_m; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth2 $wl value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']))"
done
PVB200_GATHER_DEPTH1=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth1 slowfast value %.1f ms/step %.3f' % (d['value'], d['ms_per_step']))"
