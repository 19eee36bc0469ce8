#!/bin/bash
# A/B/C on ONE box (boxes differ by ~7 %): A = previous build, B = this build without the shared-prefix restart, C = this build
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; tag=${1:-r02j}
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_scenes.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -4 gpurun_out/${tag}_pytest.txt
run() {  # name workload env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu-baseline 2> gpurun_out/${tag}_${name}_$wl.err | tail -1 > gpurun_out/${tag}_${name}_$wl.json
  python -c "
import json; d=json.load(open('gpurun_out/${tag}_${name}_$wl.json')); r=d.get('roofline',{}); print('$name $wl', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value']), 'fine launch', round(r.get('launch_ms',0),4), 'frac', round(r.get('frac',0),4), 'fast_identical', d['parity'].get('fast_identical_pixels'), d['parity'].get('e2e_last_frame_identical_pixels'))" || tail -5 gpurun_out/${tag}_${name}_$wl.err
}
for wl in c3_ico8192_4k c5_flythrough_ico8192 c2_sdf2048_1080p c1_dragon_720p; do
  [ $wl = c3_ico8192_4k ] && run A $wl PYSVO_LIB=sparse-voxel-octrees_b200/build/libsvo_base.so
  run B $wl SVO_NO_PREFIX_RESTART=1
  run C $wl SVO_DUMMY=1
done
