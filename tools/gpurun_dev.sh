#!/bin/bash
# Development wrapper around gpurun: leaves the 349 MB 8192^3 sidecar out of the snapshot for this one
# call (it costs 3-5 minutes of push time), and restores .gpurunignore afterwards. The committed
# .gpurunignore never excludes it, so round-end runs get the full default workload.
cd "$(dirname "$0")/.."
cp .gpurunignore /tmp/.gpurunignore.saved
trap 'cp /tmp/.gpurunignore.saved .gpurunignore' EXIT
echo "scenes/_cache/ico8192.words.xz" >> .gpurunignore
/usr/local/graft/bin/gpurun "$@"
