python -m pytest tests -x -q -m gpu > gpurun_out/r6_tests.txt 2>&1; tail -3 gpurun_out/r6_tests.txt
python bench.py > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; tail -c 600 gpurun_out/r6_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r6_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'step_issue', (d['roofline'].get('step_issue') or {}).get('frac'), 'e2e', d['e2e']['value'])
print('c4', d['c4']['value'], 'c4 e2e', {k: v for k, v in d['c4'].get('e2e', {}).items() if k in ('value', 'd2h_gbs', 'seconds')}, (d['c4'].get('e2e', {}).get('float64') or {}))
PY
