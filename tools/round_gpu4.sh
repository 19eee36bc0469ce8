#!/bin/bash
cd "$(dirname "$0")/.."
tag=${1:-r01e}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 300 2>gpurun_out/${tag}_c3_stderr.log | tail -1 > gpurun_out/${tag}_c3.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3.json')); print('default', d['config']['workload'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${tag}_c3_reference.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3_reference.json')); print('reference arm', d['value'], d['cpu_baseline'])"
SVO_L2_PERSIST_MB=79 python bench.py --steps 300 --no-cpu-baseline 2>gpurun_out/${tag}_l2_c3.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 with L2 window 79 MB', round(d['value']), d['ms_per_step'], d['parity'])" | tee gpurun_out/${tag}_l2_c3.txt
python tools/exp_rebuild.py ico8192 | tee gpurun_out/${tag}_rebuild_ico8192.txt
for w in c1_dragon_720p c2_sdf2048_1080p c4_ao_sdf2048 c5_flythrough_ico8192; do
  python bench.py --workload $w --steps 200 2>gpurun_out/${tag}_${w}_stderr.log | tail -1 > gpurun_out/${tag}_${w}.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${w}.json')); print('$w', round(d['value']), d['unit'], d['ms_per_step'], 'e2e', round(d['e2e']['value']), d.get('parity'), 'frac', d['roofline']['frac'], d['cpu_baseline']['value'])"
done
