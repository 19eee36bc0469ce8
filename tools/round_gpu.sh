#!/bin/bash
# One GPU-box pass of the round's evidence: parity tests, default bench (+ reference arm), load timing,
# ncu launch list and one full capture of the fine pass on the default workload. Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
tag=${1:-r01c}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 300 2>gpurun_out/${tag}_c3_stderr.log | tail -1 > gpurun_out/${tag}_c3.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3.json')); print('default', d['config']['workload'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${tag}_c3_reference.json
tail -c 600 gpurun_out/${tag}_c3_reference.json; echo
python tools/bench_load.py ico8192 --threads 1,4,16 2>gpurun_out/${tag}_load_stderr.log | tail -1 > gpurun_out/${tag}_load_ico8192.json
cat gpurun_out/${tag}_load_ico8192.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_c3_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_b_under_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:finePass -s 3 -c 1 -f -o gpurun_out/${tag}_fine_c3_4k python tools/profile_frames.py --workload c3_ico8192_4k --frames 6 2>&1 | tail -2
