#!/bin/bash
# Builds libmmdx.so (C ABI, include/mmdx.h) for sm_100a only.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libmmdx.so
stale=0
for f in engine.cu t5_decoder.cu tokenizer.cpp *.cuh ../../include/mmdx.h build.sh; do
  if [ ! -e "$OUT" ] || [ "$f" -nt "$OUT" ]; then stale=1; fi
done
if [ "$stale" = "0" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "libmmdx.so up to date"; exit 0
fi
# host-only translation units (no device code) are compiled by g++ directly and only when they changed
if [ ! -e tokenizer.o ] || [ tokenizer.cpp -nt tokenizer.o ] || [ ../../include/mmdx.h -nt tokenizer.o ]; then
  g++ -O2 -std=c++17 -fPIC -pthread -c tokenizer.cpp -o tokenizer.o
fi
if [ ! -e engine.o ] || [ "${FORCE:-0}" = "1" ] || [ -n "$(find engine.cu *.cuh ../../include/mmdx.h build.sh -newer engine.o)" ]; then
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
    -Xptxas -v -Xcompiler -fPIC,-O2 -c -o engine.o engine.cu 2> build.log || { cat build.log; exit 1; }
fi
if [ ! -e t5_decoder.o ] || [ t5_decoder.cu -nt t5_decoder.o ] || [ ../../include/mmdx.h -nt t5_decoder.o ]; then
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -Xcompiler -fPIC,-O2 -c -o t5_decoder.o t5_decoder.cu 2> build_t5.log || { cat build_t5.log; exit 1; }
fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$OUT" engine.o t5_decoder.o tokenizer.o -ldl -lpthread
grep -E "error|warning" build.log | grep -v "Wno" | head -20 || true
echo "built $OUT"
