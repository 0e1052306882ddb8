#!/usr/bin/env bash
# Builds libb200attn.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="${B200_OUT:-$here/libb200attn.so}"   # B200_OUT + B200_EXTRA_FLAGS: A/B variants of the same library
objdir="${B200_OBJDIR:-$here}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a
       -Xcompiler -fPIC --expt-relaxed-constexpr
       -I"$root/include" -I"$here" ${B200_PTXAS_V:+-Xptxas -v} ${B200_TC_TRACE:+-DB200_TC_TRACE} ${B200_DEC_TRACE:+-DB200_DEC_TRACE} ${B200_TC_POLY:+-DB200_TC_POLY=$B200_TC_POLY} ${B200_TC_POLY1:+-DB200_TC_POLY1=$B200_TC_POLY1} ${B200_EXTRA_FLAGS:-})
objs=()
pids=()
for f in decode store extend extend_tc extend_tc2; do
  "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$objdir/$f.o" &
  pids+=($!)
  objs+=("$objdir/$f.o")
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$out" "${objs[@]}" -gencode arch=compute_100a,code=sm_100a
echo "built $out"
