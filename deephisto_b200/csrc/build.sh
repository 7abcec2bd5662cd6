#!/usr/bin/env bash
# Build libdeephisto_b200.so for sm_100a (B200). nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libdeephisto_b200.so"
obj="$here/_obj"
mkdir -p "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden ${DH_NVCC_EXTRA:-})
pids=()
for f in dh_dense dh_gather dh_gather_tma dh_stitch dh_stitch_binned dh_cover dh_cnn; do
  "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$obj/$f.o" &
  pids+=($!)
done
# float64 geometry: no FMA contraction, so the CPU oracle reproduces every rounding
"$NVCC" "${FLAGS[@]}" -fmad=false -c "$here/dh_region.cu" -o "$obj/dh_region.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$out" "$obj"/dh_dense.o "$obj"/dh_gather.o "$obj"/dh_gather_tma.o "$obj"/dh_stitch.o "$obj"/dh_stitch_binned.o "$obj"/dh_cover.o "$obj"/dh_cnn.o "$obj"/dh_region.o
echo "built $out"
