#!/bin/bash
# Everything this repo runs on a GPU box, as sub-commands (run under gpurun from the repo root; output under gpurun_out/):
#   tools/gpu_session.sh tests  [tag]          every -m gpu test
#   tools/gpu_session.sh bench  [tag] [N]      the default bench line at N GPUs (torchrun for N > 1, like the driver) + the reference arm at N = 1
#   tools/gpu_session.sh lines  [tag]          bench lines of the other BASELINE configs at N = 1 (c1, c2, c4, c5)
#   tools/gpu_session.sh ncu    [tag]          launch list of the default bench + ncu --set full of the fine pass (-> profiles/traffic.json entry), K1, beam pass
#   tools/gpu_session.sh multi  [tag] N        N-GPU session: multi-GPU tests, c3 and c5 under torchrun, host-leg variants, peer counters of device 1's fine pass
#   tools/gpu_session.sh ab     [tag] LIB      A/B on one box: LIB (another build of the library) against this build, with / without the shared-prefix restart
# Boxes differ by ~7 % in the same kernel's speed: compare variants only inside one call.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
what=${1:-tests}; tag=${2:-r02}; arg=${3:-1}
line() {   # file label -> one summary line
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d.get('roofline') or {}; e = d.get('e2e') or {}
    print(sys.argv[2], d['config']['workload'], 'N=%d' % d['n_gpus'], round(d['value']), 'Mrays/s', round(d['ms_per_step'], 4), 'ms; e2e', round(e.get('value', 0)),
          'frac', round(r.get('frac', 0), 4), 'launch_ms', round(r.get('launch_ms', 0), 4), 'parity', d.get('parity'), 'clocks', (d.get('clocks') or {}).get('sm_mhz'))
except Exception as ex:
    print(sys.argv[2], 'FAILED', ex)
PY
}
bench() {  # out-file N args...
  out=$1; n=$2; shift 2
  if [ "$n" -gt 1 ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" 2> ${out%.json}.err | tail -1 > $out
  else
    timeout 900 python bench.py "$@" 2> ${out%.json}.err | tail -1 > $out
  fi
}
case $what in
tests)
  timeout 1800 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.txt; tail -16 gpurun_out/${tag}_pytest.txt ;;
bench)
  bench gpurun_out/${tag}_bench_c3_n$arg.json $arg; line gpurun_out/${tag}_bench_c3_n$arg.json own
  if [ "$arg" = 1 ]; then timeout 600 python bench.py --impl reference 2> gpurun_out/${tag}_reference.err | tail -1 > gpurun_out/${tag}_c3_reference.json; cat gpurun_out/${tag}_c3_reference.json | cut -c1-600; fi ;;
lines)
  for wl in c1_dragon_720p c2_sdf2048_1080p c5_flythrough_ico8192; do bench gpurun_out/${tag}_$wl.json 1 --workload $wl --no-cpu-baseline; line gpurun_out/${tag}_$wl.json own; done
  bench gpurun_out/${tag}_c4_ao_sdf2048.json 1 --workload c4_ao_sdf2048 --steps 20 --warmup 3; line gpurun_out/${tag}_c4_ao_sdf2048.json own ;;
ncu)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_list.log 2>&1
  for k in finePassKernel tilePrefixKernel coarsePassKernel classifyTilesKernel; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/${tag}_${k}_c3 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_$k.log 2>&1
    python tools/ncu_summary.py gpurun_out/${tag}_${k}_c3.ncu-rep --json gpurun_out/${tag}_${k}_c3.json $([ $k = finePassKernel ] && echo --traffic c3_ico8192_4k) > gpurun_out/${tag}_${k}_c3.txt 2>&1
    head -12 gpurun_out/${tag}_${k}_c3.txt
  done
  ncu --set full --clock-control none --import-source on -k regex:raymarchBatchRefill -s 4 -c 1 -f -o gpurun_out/${tag}_k1_refill_c4 python bench.py --workload c4_ao_sdf2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_k1.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_k1_refill_c4.ncu-rep --json gpurun_out/${tag}_k1_refill_c4.json > gpurun_out/${tag}_k1_refill_c4.txt 2>&1; head -12 gpurun_out/${tag}_k1_refill_c4.txt
  cp profiles/traffic.json gpurun_out/${tag}_traffic.json ;;
multi)
  N=$arg
  nvidia-smi topo -m > gpurun_out/${tag}_topo_n$N.txt 2>&1
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${tag}_pytest_multi_n$N.txt 2>&1; tail -3 gpurun_out/${tag}_pytest_multi_n$N.txt
  for wl in c3_ico8192_4k c5_flythrough_ico8192; do bench gpurun_out/${tag}_${wl}_n$N.json $N --workload $wl --no-cpu-baseline; line gpurun_out/${tag}_${wl}_n$N.json torchrun; done
  bench gpurun_out/${tag}_c3_n1_same_box.json 1 --no-cpu-baseline; line gpurun_out/${tag}_c3_n1_same_box.json same-box
  SVO_PREP_ON_LANE=1 timeout 300 python bench.py --gpus $N --no-cpu-baseline 2> gpurun_out/${tag}_preplane.err | tail -1 > gpurun_out/${tag}_c3_n${N}_prep_on_lane.json; line gpurun_out/${tag}_c3_n${N}_prep_on_lane.json prep-on-lane
  for lanes in 6 8; do SVO_MULTI_LANES=$lanes timeout 300 python bench.py --gpus $N --no-cpu-baseline 2> gpurun_out/${tag}_lanes$lanes.err | tail -1 > gpurun_out/${tag}_c3_n${N}_lanes$lanes.json; line gpurun_out/${tag}_c3_n${N}_lanes$lanes.json lanes$lanes; done
  timeout 300 ncu --devices 1 --set full --clock-control none -k regex:finePassKernel -s 30 -c 2 -f -o gpurun_out/${tag}_c5_fine_dev1_n$N \
      python bench.py --gpus $N --workload c5_flythrough_ico8192 --steps 60 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_n$N.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_c5_fine_dev1_n$N.ncu-rep --json gpurun_out/${tag}_c5_fine_dev1_n$N.json > gpurun_out/${tag}_c5_fine_dev1_n$N.txt 2>&1
  grep -i "^void\|nvl.x\|peer.*write\|duration\|hit_rate" gpurun_out/${tag}_c5_fine_dev1_n$N.txt | head -20 ;;
ab)
  for wl in c3_ico8192_4k c5_flythrough_ico8192; do
    PYSVO_LIB=$arg bench gpurun_out/${tag}_A_$wl.json 1 --workload $wl --no-cpu-baseline; line gpurun_out/${tag}_A_$wl.json "A($arg)"
    SVO_NO_PREFIX_RESTART=1 bench gpurun_out/${tag}_B_$wl.json 1 --workload $wl --no-cpu-baseline; line gpurun_out/${tag}_B_$wl.json "B(no restart)"
    bench gpurun_out/${tag}_C_$wl.json 1 --workload $wl --no-cpu-baseline; line gpurun_out/${tag}_C_$wl.json "C(this build)"
  done ;;
*) echo "unknown sub-command $what"; exit 2 ;;
esac
