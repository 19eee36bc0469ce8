#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; tag=${1:-r02f}
for devs in 0 0,0 0,0,0,0; do
  SVO_BENCH_DEVICES=$devs timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2> gpurun_out/${tag}_devs_$devs.err | tail -1 > gpurun_out/${tag}_devs_$devs.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_devs_$devs.json')); print('devices $devs', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms', d['config']['timed_region']['round_ms_min_med_max'])" || tail -20 gpurun_out/${tag}_devs_$devs.err
done
