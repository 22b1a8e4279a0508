#!/usr/bin/env bash
# Rebuild the library on the GPU box with different compile-time knobs and time the headline leg (development aid).
out=${1:-gpurun_out/variants.txt}
shift
for v in "$@"; do
  NVCC_EXTRA="$v" bash geograypher_b200/csrc/build.sh > /dev/null 2>&1
  echo "== $v" >> $out
  python bench.py --steps 4 --warmup 3 --skip c3,c4,c5,e2e,cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['pixel_sum']['value']), d['stage_ms'])" >> $out
done
NVCC_EXTRA="" bash geograypher_b200/csrc/build.sh > /dev/null 2>&1
