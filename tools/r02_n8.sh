#!/bin/bash
# The 8-GPU session: c3 and the c5 fly-through at N ranks under torchrun (the driver's launch), host-leg stripe widths,
# peer-traffic / L2 counters of one non-zero device's fine pass.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; N=${1:-8}; tag=${2:-r02h}
nvidia-smi topo -m > gpurun_out/${tag}_topo_n$N.txt 2>&1
for wl in c3_ico8192_4k c5_flythrough_ico8192; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
      --workload $wl --steps 300 --warmup 10 --no-cpu-baseline 2> gpurun_out/${tag}_${wl}_n$N.err | tail -1 > gpurun_out/${tag}_${wl}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${wl}_n$N.json')); print('N=$N', d['config']['workload'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms', d['parity'], d['config']['timed_region'])" || tail -20 gpurun_out/${tag}_${wl}_n$N.err
done
for run in 15 60; do
  SVO_MULTI_HOST_RUN=$run timeout 300 python bench.py --gpus $N --steps 300 --warmup 10 --no-cpu-baseline 2> gpurun_out/${tag}_run$run.err | tail -1 > gpurun_out/${tag}_c3_n${N}_run$run.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3_n${N}_run$run.json')); print('plain process N=$N host stripes $run: device', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms =', round(d['e2e']['d2h_bytes_per_step']/d['e2e']['ms_per_step']/1e6,1), 'GB/s')" || tail -5 gpurun_out/${tag}_run$run.err
done
timeout 300 ncu --devices 1 --set full --clock-control none -k regex:finePassKernel -s 30 -c 2 -f -o gpurun_out/${tag}_c5_fine_dev1_n$N \
    python bench.py --gpus $N --workload c5_flythrough_ico8192 --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_n$N.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_c5_fine_dev1_n$N.ncu-rep --json gpurun_out/${tag}_c5_fine_dev1_n$N.json > gpurun_out/${tag}_c5_fine_dev1_n$N.txt 2>&1
grep -i "^void\|nvl.x\|peer.*write_lookup\|duration\|hit_rate" gpurun_out/${tag}_c5_fine_dev1_n$N.txt | head -20
