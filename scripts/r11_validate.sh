python -m pytest tests -x -q -m gpu > gpurun_out/r11_tests.txt 2>&1; tail -3 gpurun_out/r11_tests.txt
python scripts/prof_c5_e2e.py > gpurun_out/r11_c5_e2e.txt 2> gpurun_out/r11_c5_e2e.err; cat gpurun_out/r11_c5_e2e.txt; tail -3 gpurun_out/r11_c5_e2e.err
python bench.py > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err; tail -c 500 gpurun_out/r11_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r11_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'step', d['roofline']['step_issue']['frac'], 'e2e', d['e2e']['value'])
print('c4', d['c4']['value'], d['c4']['e2e'].get('value'), d['c4']['e2e'].get('float64', {}).get('value'), d['c4']['e2e'].get('error'))
print('c5', d['c5']['value'], d['c5']['e2e'].get('value'), d['c5']['e2e'].get('parity_with_resident_leg'), d['c5']['e2e'].get('error'))
PY
