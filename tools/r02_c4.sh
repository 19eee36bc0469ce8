#!/bin/bash
# c4 (incoherent AO rays through the batch API): tests of the refill kernel / own counting sort, then the bench line with and without lane refill
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; tag=${1:-r02b}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scenes.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -5 gpurun_out/${tag}_pytest.txt
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "c4 or c2" >> gpurun_out/${tag}_pytest.txt 2>&1; tail -3 gpurun_out/${tag}_pytest.txt
for mode in refill norefill; do
  if [ $mode = norefill ]; then export SVO_BENCH_NO_LANE_REFILL=1; else unset SVO_BENCH_NO_LANE_REFILL; fi
  timeout 600 python bench.py --workload c4_ao_sdf2048 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_c4_$mode.err | tail -1 > gpurun_out/${tag}_c4_$mode.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_c4_$mode.json')); print('$mode', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), d['parity'])" || tail -5 gpurun_out/${tag}_c4_$mode.err
done
