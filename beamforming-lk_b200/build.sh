#!/bin/bash
# Builds beamforming-lk_b200/libbflk.so (CUDA kernels + C ABI) for sm_100a, in-tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off ${EXTRA_NVCC_FLAGS:-}"
mkdir -p build
objs=()
for f in csrc/bflk_api.cu csrc/tables.cu csrc/das_generic.cu csrc/das_tile.cu csrc/das_bcast.cu csrc/post.cu; do
  o=build/$(basename "${f%.cu}").o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ csrc/bflk_internal.h -nt "$o" ] || [ csrc/das_common.cuh -nt "$o" ] || [ csrc/das_tile_asm.inc -nt "$o" ] || [ csrc/das_tile_fast_asm.inc -nt "$o" ] || [ ../include/bflk.h -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
  fi
  objs+=("$o")
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libbflk.so "${objs[@]}"
echo "built $(pwd)/libbflk.so"
