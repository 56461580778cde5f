#!/bin/sh
# Cross-compile the CPU-only forward runtime with the Android NDK (see README.md).
set -e
NDK=${1:?usage: build_android.sh <ndk-root> [abi] [platform]}
ABI=${2:-arm64-v8a}
PLATFORM=${3:-android-24}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${OUT:-$HERE/../../../../build/android-$ABI}
cmake -S "$HERE/.." -B "$OUT" \
  -DCMAKE_TOOLCHAIN_FILE="$NDK/build/cmake/android.toolchain.cmake" \
  -DANDROID_ABI="$ABI" -DANDROID_PLATFORM="$PLATFORM" -DANDROID_ARM_NEON=ON \
  -DCMAKE_BUILD_TYPE=Release -DZNICZ_WITH_CUDA=OFF -DZNICZ_OPENMP=ON
cmake --build "$OUT" -j
echo "built: $OUT"
