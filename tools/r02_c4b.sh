#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; tag=${1:-r02d}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scenes.py -m gpu -x -q -k "batch or ambient" > gpurun_out/${tag}_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_pytest.txt
for idle in 4 8 12 16 20; do
  SVO_REFILL_IDLE=$idle timeout 600 python bench.py --workload c4_ao_sdf2048 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_c4_idle$idle.err | tail -1 > gpurun_out/${tag}_c4_idle$idle.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_c4_idle$idle.json')); print('idle $idle', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), d['parity'])" || tail -5 gpurun_out/${tag}_c4_idle$idle.err
done
