mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_bench_contract.py -q -m gpu --timeout 800 -x > gpurun_out/tests_r03e.log 2>&1; echo tests rc=$?; tail -3 gpurun_out/tests_r03e.log
python bench.py --steps 20 --warmup 5 --no-eager --no-cpu-baseline > gpurun_out/bench_r03e.log 2>&1; python - <<'PY'
import json
for line in open('gpurun_out/bench_r03e.log'):
    if line.startswith('{'):
        d=json.loads(line)
        print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['fp32_host_tensors']['value'])
        print(d['configs']['config4_single_style'])
        print({k:v.get('value') for k,v in d['configs'].items()})
PY
