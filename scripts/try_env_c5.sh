#!/usr/bin/env bash
# Like try_env.sh for the c5 leg (20 M faces, 8192x5460 rig views, vote mode)   (development aid)
out=${1:-gpurun_out/env_c5.txt}
shift
for v in "$@"; do
  echo "== $v" >> $out
  env $v python bench.py --steps 2 --warmup 1 --skip c3,c4,e2e,cpu,pixel_sum 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['c5']['value']), d['c5']['stage_ms'])" >> $out
done
