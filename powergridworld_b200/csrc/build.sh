#!/usr/bin/env bash
# Build libpgw_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libpgw_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
mkdir -p "$HERE/build"
# component kernels: no FMA contraction, so float64 rounds exactly like the reference
"$NVCC" $COMMON -fmad=false ${PTXAS_V:+-Xptxas -v} -c "$HERE/components.cu" -o "$HERE/build/components.o"
"$NVCC" $COMMON ${PTXAS_V:+-Xptxas -v} -c "$HERE/powerflow.cu" -o "$HERE/build/powerflow.o"
"$NVCC" $COMMON ${PTXAS_V:+-Xptxas -v} -c "$HERE/powerflow_tc.cu" -o "$HERE/build/powerflow_tc.o"
"$NVCC" $COMMON ${PTXAS_V:+-Xptxas -v} -c "$HERE/powerflow_tc2.cu" -o "$HERE/build/powerflow_tc2.o"
"$NVCC" $COMMON -c "$HERE/api.cu" -o "$HERE/build/api.o"
"$NVCC" -shared $ARCH -o "$OUT" "$HERE/build/components.o" "$HERE/build/powerflow.o" "$HERE/build/powerflow_tc.o" "$HERE/build/powerflow_tc2.o" "$HERE/build/api.o" -lcudart
echo "built $OUT"
