#!/usr/bin/env bash
# Like try_variants.sh, but every variant first has to pass the parity tests (development aid).
out=${1:-gpurun_out/variants.txt}
shift
for v in "$@"; do
  NVCC_EXTRA="$v" bash geograypher_b200/csrc/build.sh > /dev/null 2>&1
  echo "== $v" >> $out
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -1 >> $out
  python bench.py --steps 4 --warmup 3 --skip c3,c5,e2e,cpu,pixel_sum 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['c4']['value']), d['stage_ms'])" >> $out
done
NVCC_EXTRA="" bash geograypher_b200/csrc/build.sh > /dev/null 2>&1
