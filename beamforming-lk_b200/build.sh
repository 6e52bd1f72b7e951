#!/bin/bash
# Builds beamforming-lk_b200/libbflk.so (CUDA kernels + C ABI) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off ${EXTRA_NVCC_FLAGS:-}"
mkdir -p build
objs=()
pids=()
for f in csrc/*.cu; do
  o=build/$(basename "${f%.cu}").o
  stale=0
  [ -f "$o" ] || stale=1
  for dep in "$f" csrc/*.h csrc/*.cuh csrc/*.inc ../include/bflk.h build.sh; do
    [ "$dep" -nt "$o" ] && stale=1
  done
  if [ $stale = 1 ]; then
    rm -f "$o"                      # a failed compile must not leave an older object for the link step
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
  objs+=("$o")
done
for p in "${pids[@]:-}"; do
  [ -n "$p" ] && { wait "$p" || { echo "build.sh: a compile failed" >&2; exit 1; }; }
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libbflk.so "${objs[@]}" -ldl
echo "built $(pwd)/libbflk.so"
