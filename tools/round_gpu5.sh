#!/bin/bash
# End-of-session verification on the GPU box: every GPU test, smoke(), the default bench line and the reference arm,
# and a memcheck pass over the kernels added in this session.
cd "$(dirname "$0")/.."
tag=${1:-r01h}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 300 2>gpurun_out/${tag}_c3_stderr.log | tail -1 > gpurun_out/${tag}_c3.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3.json')); print('default', d['config']['workload'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['parity'], 'frac', d['roofline']['frac'], d['roofline']['launch_ms'], d['cpu_baseline'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${tag}_c3_reference.json
python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3_reference.json')); print('reference arm', d['value'], d['cpu_baseline'])"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_build.py -x -q -m gpu -k "owned_tiles or viewer or build_from_ply" > gpurun_out/${tag}_sanitizer.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/${tag}_sanitizer.txt
