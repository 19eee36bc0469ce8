#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the multi-GPU tests, c3 and c5 at N ranks under torchrun (the driver's launch), NVLink counters.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; N=${1:-2}; tag=${2:-r02e}
nvidia-smi topo -m > gpurun_out/${tag}_topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${tag}_pytest_multi_n$N.txt 2>&1; tail -3 gpurun_out/${tag}_pytest_multi_n$N.txt
for wl in c3_ico8192_4k c5_flythrough_ico8192; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
      --workload $wl --steps 300 --warmup 10 --no-cpu-baseline 2> gpurun_out/${tag}_${wl}_n$N.err | tail -1 > gpurun_out/${tag}_${wl}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${wl}_n$N.json')); print('N=$N', d['config']['workload'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms', d['parity'], d['config']['timed_region'])" || tail -20 gpurun_out/${tag}_${wl}_n$N.err
done
# NVLink / peer traffic of the fused gather: the fine pass of a non-zero device (single process, so ncu sees every device)
timeout 600 ncu --set full --clock-control none -k regex:finePassKernel -s 40 -c $N -f -o gpurun_out/${tag}_fine_peer_n$N \
    python bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_peer_n$N.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_fine_peer_n$N.ncu-rep --json gpurun_out/${tag}_fine_peer_n$N.json > gpurun_out/${tag}_fine_peer_n$N.txt 2>&1
grep -i "kernel\|nvl\|peer\|duration\|sysmem" gpurun_out/${tag}_fine_peer_n$N.txt | head -40
