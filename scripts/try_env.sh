#!/usr/bin/env bash
# Time the headline leg under different run-time knobs (environment variables), e.g.
#   bash scripts/try_env.sh gpurun_out/env.txt "GG_FILL_CTAS=2" "GG_FILL_CTAS=2 GG_SETUP_CTAS=2"   (development aid)
out=${1:-gpurun_out/env.txt}
shift
for v in "$@"; do
  echo "== $v" >> $out
  env $v python bench.py --steps 4 --warmup 3 --skip c3,c4,c5,e2e,cpu,pixel_sum 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['stage_ms'])" >> $out
done
