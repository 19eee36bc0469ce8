#!/bin/bash
# one-GPU experiment: swap in library variants and time the fine pass
cd "$(dirname "$0")/.."
cp sparse-voxel-octrees_b200/libsvo_b200.so /tmp/orig.so
for v in "$@"; do
  cp sparse-voxel-octrees_b200/libsvo_b200_$v.so sparse-voxel-octrees_b200/libsvo_b200.so
  echo "variant $v"
  python tools/exp_interleave.py c2_sdf2048_4k | grep "world=1:"
  python tools/exp_interleave.py c1_dragon_720p | grep "world=1:"
done
cp /tmp/orig.so sparse-voxel-octrees_b200/libsvo_b200.so
