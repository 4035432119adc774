#!/bin/bash
# Builds libmmdx.so (C ABI, include/mmdx.h) for sm_100a only.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libmmdx.so
if [ "$OUT" -nt engine.cu ] && [ "$OUT" -nt kernels.cuh ] && [ "$OUT" -nt gemm_tcgen05.cuh ] && [ "$OUT" -nt ptx.cuh ] \
   && [ "$OUT" -nt ../../include/mmdx.h ] && [ "${FORCE:-0}" != "1" ]; then
  echo "libmmdx.so up to date"; exit 0
fi
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
  -Xptxas -v -shared -Xcompiler -fPIC,-O2 -o "$OUT" engine.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $OUT"
