#!/bin/bash
# Build crazyflie_nmpc_b200/variants/libcfnmpc_<name>.so from the current tree with the given patches applied to a scratch
# copy of csrc/ (A/B experiments on one GPU box; the library under test is selected with CFNMPC_LIB).
# Usage: tools/build_variant.sh <name> [patch ...] [-- extra nvcc flags]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
TMP=$(mktemp -d)
mkdir -p "$TMP/crazyflie_nmpc_b200" "$ROOT/crazyflie_nmpc_b200/variants"
cp -r "$ROOT/crazyflie_nmpc_b200/csrc" "$TMP/crazyflie_nmpc_b200/"
cp -r "$ROOT/include" "$TMP/"
extra=()
while [ $# -gt 0 ]; do
  if [ "$1" == "--" ]; then shift; extra=("$@"); break; fi
  (cd "$TMP" && patch -s -p1 < "$ROOT/$1")
  shift
done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "${extra[@]}" \
  -I "$TMP/include" "$TMP/crazyflie_nmpc_b200/csrc/cfnmpc_api.cu" "$TMP/crazyflie_nmpc_b200/csrc/acados_shim.cpp" "$TMP/crazyflie_nmpc_b200/csrc/cfnmpc_multi.cpp" \
  -o "$ROOT/crazyflie_nmpc_b200/variants/libcfnmpc_$name.so"
rm -rf "$TMP"
echo "built variants/libcfnmpc_$name.so"
