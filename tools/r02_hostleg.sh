#!/bin/bash
# host leg of a multi-GPU frame: copy engine (2-D copies per stripe) vs zero-copy kernel, stripe widths
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; N=${1:-2}; tag=${2:-r02g}
for mode in engine kernel; do for run in 15 30 60; do
  SVO_MULTI_HOST_COPY=$mode SVO_MULTI_HOST_RUN=$run timeout 600 python bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline 2> gpurun_out/${tag}_$mode_$run.err | tail -1 > gpurun_out/${tag}_${mode}_$run.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${mode}_$run.json')); print('N=$N $mode run $run: device', round(d['value']), 'Mrays/s; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms =', round(d['e2e']['d2h_bytes_per_step']/d['e2e']['ms_per_step']/1e6,1), 'GB/s', d['parity'].get('e2e_host_frame_identical_to_single_rank'))" || tail -20 gpurun_out/${tag}_$mode_$run.err
done; done
