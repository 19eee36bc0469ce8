cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for wl in c3_ico8192_4k c1_dragon_720p c2_sdf2048_1080p; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r02q_$wl.err | tail -1 > gpurun_out/r02q_$wl.json
  python -c "
import json; d=json.load(open('gpurun_out/r02q_$wl.json')); g=d['e2e_grey8a8']; print('$wl', round(d['value']), 'Mrays/s; e2e rgba', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms; e2e grey8a8', round(g['value']), round(g['ms_per_step'],4), 'ms', g['expands_to_the_rgba_frame'])" || tail -5 gpurun_out/r02q_$wl.err
done
