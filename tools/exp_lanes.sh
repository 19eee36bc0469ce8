#!/bin/bash
# experiment: frames in flight per rank vs multi-GPU throughput (workload without the big sidecar)
cd "$(dirname "$0")/.."
N=${1:-4}
for L in ${LANES:-2 3 4}; do
  SVO_BENCH_LANES=$L python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c2_sdf2048_4k --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=$N lanes=$L', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']))"
done
