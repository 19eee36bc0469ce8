#!/bin/bash
# Round-2 GPU session script (run under gpurun from the repo root): tests, the default bench line, ncu captures.
#   tools/r02_gpu.sh tests|bench|ncu|all [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${1:-all}; tag=${2:-r02a}
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.txt
  tail -25 gpurun_out/${tag}_pytest.txt
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 900 python bench.py 2> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench_c3_n1.json
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${tag}_bench_c3_n1.json'))
    print('c3 N=1', round(d['value']), 'Mrays/s', round(d['ms_per_step'], 4), 'ms; e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 4),
          'clocks', d['clocks'], 'parity', d['parity'], 'cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${tag}_bench.err').read()[-3000:])
PY
fi
if [ "$what" = ncu ] || [ "$what" = all ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv \
      python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_list.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:finePassKernel -s 12 -c 1 -f -o gpurun_out/${tag}_fine_c3 \
      python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_fine.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_fine_c3.ncu-rep --json gpurun_out/${tag}_fine_c3.json > gpurun_out/${tag}_fine_c3.txt 2>&1
  head -60 gpurun_out/${tag}_fine_c3.txt
fi
