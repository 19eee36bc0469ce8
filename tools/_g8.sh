cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --gpus 8 --steps 300 --warmup 5 --no-cpu-baseline 2> gpurun_out/r02s_k300.err | tail -1 > gpurun_out/r02s_c3_n8_steps300.json
python -c "
import json; d=json.load(open('gpurun_out/r02s_c3_n8_steps300.json')); g=d['e2e_grey8a8']; print('N=8 steps 300:', round(d['value']), 'Mrays/s device; e2e rgba', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms; e2e grey8a8', round(g['value']), round(g['ms_per_step'],4), 'ms', g['expands_to_the_rgba_frame'])" || tail -5 gpurun_out/r02s_k300.err
