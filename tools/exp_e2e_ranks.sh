#!/bin/bash
# experiment: host-visible frames at N > 1 -- every rank ships its own tiles (default) vs rank 0 copies the frame
cd "$(dirname "$0")/.."
N=${1:-2}
W=${2:-c2_sdf2048_4k}
for MODE in per-rank rank0; do
  if [ $MODE = rank0 ]; then export SVO_BENCH_E2E_VIA_RANK0=1; else unset SVO_BENCH_E2E_VIA_RANK0; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $W --steps 300 --warmup 10 --no-cpu-baseline 2>gpurun_out/e2e_ranks_${MODE}_n$N.err | tail -1 > gpurun_out/e2e_ranks_${MODE}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/e2e_ranks_${MODE}_n$N.json')); print('N=$N $MODE', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms', d['parity'])" || tail -5 gpurun_out/e2e_ranks_${MODE}_n$N.err
done
