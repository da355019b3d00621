#!/bin/bash
# Assembles the Swift package that runs the reference's XCTest files against libairwave_cuda.so.
#   usage: assemble_package.sh <reference checkout> <output dir>
# Needs: a Swift >= 5.9 toolchain (Linux is fine: the shim uses no Apple framework), an sm_100 GPU, and the built library
# (python -m airwave_b200.build).  Then:  cd <output dir> && AIRWAVE_CUDA_LIB_DIR=<repo>/airwave_b200/lib swift test
set -euo pipefail
REF=${1:?reference checkout}; OUT=${2:?output dir}
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
mkdir -p "$OUT/Sources/CAirwaveCUDA" "$OUT/Sources/AirwaveCUDA" "$OUT/Tests/AirwaveCUDATests"
cp "$HERE/Package.swift" "$OUT/Package.swift"
cp "$ROOT/include/module.modulemap" "$ROOT/include/airwave_cuda.h" "$OUT/Sources/CAirwaveCUDA/"
cp "$HERE/AirwaveCUDA.swift" "$OUT/Sources/AirwaveCUDA/"
for t in ConvolutionEngineTests RealtimeAudioProcessorTests; do
  # the reference's tests, unchanged except for the module they import
  sed -e 's/@testable import Airwave$/@testable import AirwaveCUDA/' "$REF/AirwaveTests/$t.swift" > "$OUT/Tests/AirwaveCUDATests/$t.swift"
done
echo "package assembled in $OUT"
