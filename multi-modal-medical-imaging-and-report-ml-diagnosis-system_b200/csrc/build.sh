#!/bin/bash
# Builds libmmdx.so (C ABI, include/mmdx.h) for sm_100a only.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libmmdx.so
stale=0
for f in engine.cu *.cuh ../../include/mmdx.h build.sh; do
  if [ ! -e "$OUT" ] || [ "$f" -nt "$OUT" ]; then stale=1; fi
done
if [ "$stale" = "0" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "libmmdx.so up to date"; exit 0
fi
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
  -Xptxas -v -shared -Xcompiler -fPIC,-O2 -o "$OUT" engine.cu -ldl 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $OUT"
