#!/bin/bash
# Builds libat_b200.so for sm_100a, in-tree (audio-tokens_b200/at_b200/libat_b200.so).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${AT_OUT:-$HERE/../at_b200/libat_b200.so}"   # AT_OUT / AT_OBJ / AT_EXTRA_FLAGS: experiment builds beside the product library
OBJ="${AT_OBJ:-$HERE/obj}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++
       -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr -Xptxas -v ${AT_EXTRA_FLAGS:-})
mkdir -p "$OBJ"
pids=()
for f in at_util at_kmeans at_mel at_assign_tc at_resample at_peer at_tokens at_conv; do
  ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$OBJ/$f.o" > "$OBJ/$f.log" 2>&1 || { cat "$OBJ/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "$OBJ"/at_util.o "$OBJ"/at_kmeans.o "$OBJ"/at_mel.o "$OBJ"/at_assign_tc.o "$OBJ"/at_resample.o "$OBJ"/at_peer.o "$OBJ"/at_tokens.o "$OBJ"/at_conv.o -lcudart
echo "built $OUT"
