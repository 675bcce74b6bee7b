#!/bin/bash
# A/B builds of the library with extra nvcc flags: scripts/build_variant.sh <name> <flags...> -> hicdiff_b200/lib/variants/lib<name>.so
# (selected at run time with HICDIFF_B200_LIB=<path>; the variants are git-ignored like the main library)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/hicdiff_b200/build/variant_$NAME
mkdir -p $OBJ $ROOT/hicdiff_b200/lib/variants
pids=()
for f in $ROOT/hicdiff_b200/csrc/*.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" -c $f -o $OBJ/$(basename $f .cu).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/hicdiff_b200/lib/variants/lib$NAME.so $OBJ/*.o
echo $ROOT/hicdiff_b200/lib/variants/lib$NAME.so
