#!/usr/bin/env bash
# Builds libgeograypher_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libgeograypher_b200.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2
       --fmad=true -Xptxas -v -ccbin /usr/bin/g++ ${NVCC_EXTRA:-})
mkdir -p build
for f in gg_api gg_raster gg_aggregate gg_warp gg_polygons; do
  "$NVCC" "${FLAGS[@]}" -c $f.cu -o build/$f.o 2> build/$f.ptxas.log || { cat build/$f.ptxas.log; exit 1; }
done
"$NVCC" -shared -o "$OUT" build/gg_api.o build/gg_raster.o build/gg_aggregate.o build/gg_warp.o build/gg_polygons.o -lcudart -lpthread -ccbin /usr/bin/g++
echo "built $OUT"
