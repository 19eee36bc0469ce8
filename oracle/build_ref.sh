#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref/libsvo_ref.so: the reference's own sources, compiled where
# they lie under $SVO_REFERENCE (default /root/reference), plus the C-ABI
# harness oracle/ref_harness.cpp. Outputs go to oracle/_ref/ only (git-ignored,
# but it travels to the GPU box with the gpurun snapshot). The reference's own
# build system is not used (its CMakeLists.txt requires SDL, CMakeLists.txt:30).
#
# Flags: -O3 -DNDEBUG (the reference's release flags) and -ffp-contract=off so
# that x86 code never fuses a*b+c -- the counterpart of nvcc -fmad=false
# (SURVEY.md section 7 "Hard parts", App. E.8).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ref="${SVO_REFERENCE:-/root/reference}"
out="$here/_ref"
src="$ref/src"

if [ ! -d "$src" ]; then
    echo "build_ref.sh: $src not found; keeping any prebuilt $out/libsvo_ref.so" >&2
    exit 3
fi
mkdir -p "$out"
tmp="$(mktemp -d /tmp/svo_ref_build.XXXXXX)"
trap 'rm -rf "$tmp"' EXIT

# Main.cpp:56-62 -- "adapt this to your platform": make the three constants
# (and the aspect ratio derived from them) run-time variables. Temp file only.
sed -e 's/^static const int NumThreads = /static int NumThreads = /' \
    -e 's/^static const int GWidth  = /static int GWidth  = /' \
    -e 's/^static const int GHeight = /static int GHeight = /' \
    -e 's/^static const float AspectRatio = /static float AspectRatio = /' \
    "$src/Main.cpp" > "$tmp/Main_runtime_dims.cpp"
for sym in NumThreads GWidth GHeight AspectRatio; do
    if grep -Eq "^static const (int|float) +$sym" "$tmp/Main_runtime_dims.cpp"; then
        echo "build_ref.sh: failed to patch $sym in Main.cpp" >&2; exit 1
    fi
done

CXXFLAGS="-std=c++11 -O3 -DNDEBUG -ffp-contract=off -fPIC -pthread -w"
CFLAGS="-O3 -DNDEBUG -ffp-contract=off -fPIC -w"
objs=()
for f in VoxelOctree.cpp VoxelData.cpp PlyLoader.cpp Util.cpp Debug.cpp \
         thread/ThreadPool.cpp thread/ThreadUtils.cpp math/Mat4.cpp math/MatrixStack.cpp; do
    o="$tmp/$(echo "$f" | tr '/' '_').o"
    g++ $CXXFLAGS -I"$src" -c "$src/$f" -o "$o" &
    objs+=("$o")
done
# the interactive path (row f4): the reference's event state and frame barrier, against the shim's SDL
for f in Events.cpp ThreadBarrier.cpp; do
    o="$tmp/$f.o"
    g++ $CXXFLAGS -I"$here/ref_shim" -I"$src" -c "$src/$f" -o "$o" &
    objs+=("$o")
done
for f in third-party/lz4.c third-party/plyfile.c third-party/tribox3.c; do
    o="$tmp/$(echo "$f" | tr '/' '_').o"
    gcc $CFLAGS -I"$src" -c "$src/$f" -o "$o" &
    objs+=("$o")
done
g++ $CXXFLAGS -I"$here/ref_shim" -I"$src" -DSVO_REF_MAIN_CPP="\"$tmp/Main_runtime_dims.cpp\"" \
    -c "$here/ref_harness.cpp" -o "$tmp/ref_harness.o" &
objs+=("$tmp/ref_harness.o")
wait
for o in "${objs[@]}"; do [ -f "$o" ] || { echo "build_ref.sh: missing $o" >&2; exit 1; }; done

g++ -shared -pthread -o "$out/libsvo_ref.so" "${objs[@]}"
echo "built $out/libsvo_ref.so"
