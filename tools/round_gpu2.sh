#!/bin/bash
# Second evidence pass: full GPU suite (with the 8192^3 round trip), builder benches, the other workloads.
cd "$(dirname "$0")/.."
tag=${1:-r01c}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_build.py --rebuild ico8192 2>/dev/null | tail -1 | tee gpurun_out/${tag}_rebuild_ico8192.json
python tools/bench_build.py --rebuild sdf2048 2>/dev/null | tail -1 | tee gpurun_out/${tag}_rebuild_sdf2048.json
python tools/bench_build.py --res 1024 2>/dev/null | tail -1 | tee gpurun_out/${tag}_build_sdf1024.json
for w in c1_dragon_720p c2_sdf2048_1080p c4_ao_sdf2048 c5_flythrough_ico8192; do
  python bench.py --workload $w --steps 200 2>gpurun_out/${tag}_${w}_stderr.log | tail -1 > gpurun_out/${tag}_${w}.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${w}.json')); print('$w', round(d['value']), d['unit'], d['ms_per_step'], 'e2e', round(d['e2e']['value']), d.get('parity'), 'frac', d['roofline']['frac'], d['cpu_baseline']['value'])"
done
