#!/usr/bin/env bash
# compute-sanitizer over smoke() and the stress tests that exercise the warp-synchronous shared-memory hand-offs of
# k_raster_tiles (long tile lists, triangle soup, near-plane clipping).  Run on the GPU box:  bash scripts/sanitize.sh
set -u
OUT=${1:-gpurun_out}
TESTS="tests/test_gpu_edge_cases.py::test_long_tile_lists tests/test_gpu_parity.py::test_triangle_soup_matches_oracle tests/test_gpu_edge_cases.py::test_near_plane_clipping tests/test_gpu_parity.py::test_fused_pixel_sum_matches_numpy"
for tool in memcheck racecheck; do
  log=$OUT/r02_sanitizer_$tool.txt
  echo "# compute-sanitizer --tool $tool  (smoke() + $TESTS)" > $log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" | head -40 >> $log
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $TESTS -x -q -m gpu 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" | head -40 >> $log
done
cat $OUT/r02_sanitizer_*.txt
