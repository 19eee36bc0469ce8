#!/bin/bash
cd "$(dirname "$0")/.."
python bench.py --steps 300 2>gpurun_out/c3_final_stderr.log | tail -1 > gpurun_out/c3_final.json
python -c "
import json; d=json.load(open('gpurun_out/c3_final.json')); print('default', d['config']['workload'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'], d['parity'], d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_launches_c3_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_under_ncu_c3.log 2>&1
cp sparse-voxel-octrees_b200/libsvo_b200.so /tmp/orig.so
cp sparse-voxel-octrees_b200/libsvo_b200_r32.so sparse-voxel-octrees_b200/libsvo_b200.so
python bench.py --steps 300 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('r32', round(d['value']), d['ms_per_step'], d['roofline']['launch_ms'])"
cp /tmp/orig.so sparse-voxel-octrees_b200/libsvo_b200.so
