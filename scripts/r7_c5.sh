python bench.py --skip pixel_sum,c3,c4,cpu --steps 2 --warmup 3 > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err; tail -c 800 gpurun_out/r7_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r7_bench.json').read().strip().splitlines()[-1])
c5 = d['c5']
print('c5', c5['value'], c5['faces_observed'], c5['votes'])
print('c5 e2e', c5.get('e2e'))
PY
python -m pytest tests -x -q -m gpu -k "vote or index or Index" 2>&1 | tail -3
