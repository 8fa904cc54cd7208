#!/bin/bash
# ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles a few reference translation units UNMODIFIED, where they lie under
# /root/reference/src, against the stand-in headers of oracle/ref_stub/ into oracle/_ref/libref_units.so (git-ignored,
# travels to the GPU box with the snapshot).  -O2 -msse2 -ffp-contract=off: the flags of the oracle's parity build.
# The reference's own build system (catkin/CMake + Eigen3 + boost + OpenCV + ...) is not used and could not run here.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SOSBA_REFERENCE:-/root/reference}/src"
[ -d "$REF" ] || { echo "ref_build.sh: $REF not found" >&2; exit 2; }
mkdir -p "$HERE/_ref"
g++ -std=c++17 -O2 -msse2 -ffp-contract=off -fPIC -shared -fvisibility=hidden -w \
    -I "$HERE/ref_stub" -I "$REF" \
    "$HERE/ref_shim.cpp" "$REF/util/settings.cpp" -o "$HERE/_ref/libref_units.so"
echo "built $HERE/_ref/libref_units.so from $REF"
