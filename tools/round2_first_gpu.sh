#!/bin/bash
# The GPU measurements round 1 did not reach. Usage on the GPU box:
#   tools/round2_first_gpu.sh one      # one GPU: ncu of the two kernels added last, first-call latency of the mesh -> tree path
#   tools/round2_first_gpu.sh N        # N = 4 or 8 GPUs (gpurun --gpus N): the default workload (8192^3 @ 4K) at N ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${TAG:-r02a}
if [ "$1" = "one" ]; then
  # copyOwnedColumnsKernel into mapped host memory (what a PCIe write burst carries) and assembleTrianglesKernel
  ncu --set full --clock-control none --import-source on -k regex:copyOwnedColumns -s 4 -c 1 -f -o gpurun_out/${tag}_copy_owned \
      python tools/exp_zero_copy.py > gpurun_out/${tag}_copy_owned.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_copy_owned.ncu-rep --json gpurun_out/${tag}_copy_owned.json > gpurun_out/${tag}_copy_owned.txt 2>&1
  ncu --set full --clock-control none --import-source on -k regex:assembleTriangles -c 1 -f -o gpurun_out/${tag}_assemble \
      python tools/exp_ply.py ico2048 > gpurun_out/${tag}_assemble.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_assemble.ncu-rep --json gpurun_out/${tag}_assemble.json > gpurun_out/${tag}_assemble.txt 2>&1
  # first build of a process against the following ones (three builds per process, two processes)
  for i in 1 2; do python tools/exp_ply.py ico8192 2>&1 | grep -v "voxelizeMesh\|OctreeBuilder" | tail -12; done | tee gpurun_out/${tag}_ply_first_call.txt
else
  N=$1
  # configs[4]: the fly-through sweep on the 8192^3 tree at N ranks (ms per frame; the L2 hit rate comes from the ncu capture at N = 1)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N \
      --workload c5_flythrough_ico8192 --steps 300 --warmup 10 --no-cpu-baseline 2>gpurun_out/${tag}_c5_n$N.err | tail -1 > gpurun_out/${tag}_c5_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_c5_n$N.json')); print('N=$N', d['config']['workload'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), d['parity'])" || tail -5 gpurun_out/${tag}_c5_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
      --steps 300 --warmup 10 --no-cpu-baseline 2>gpurun_out/${tag}_c3_n$N.err | tail -1 > gpurun_out/${tag}_c3_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_c3_n$N.json')); print('N=$N', d['config']['workload'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'ms', d['parity'])" || tail -5 gpurun_out/${tag}_c3_n$N.err
fi
