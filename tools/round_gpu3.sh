#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r01d}
python bench.py --steps 300 2>gpurun_out/${tag}_c3_stderr.log | tail -1 > gpurun_out/${tag}_c3.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3.json')); print('default', d['config']['workload'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'])"
python tools/exp_rebuild.py ico8192
cp sparse-voxel-octrees_b200/libsvo_b200.so /tmp/new.so; cp tmp_so/libsvo_old.so sparse-voxel-octrees_b200/libsvo_b200.so; echo OLD
python tools/exp_rebuild.py ico8192
cp /tmp/new.so sparse-voxel-octrees_b200/libsvo_b200.so
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches_c3_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_b_under_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:finePass -s 3 -c 1 -f -o gpurun_out/${tag}_fine_c3_4k python tools/profile_frames.py --workload c3_ico8192_4k --frames 6 2>&1 | tail -1
ncu --set full --clock-control none -k regex:finePass -s 3 -c 1 -f -o gpurun_out/${tag}_fine_c5_mid python tools/profile_frames.py --workload c5_flythrough_ico8192 --frames 6 --cam 50 2>&1 | tail -1
