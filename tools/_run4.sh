set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_r2_api.py -x -q > gpurun_out/r02i_pytest.log 2>&1; tail -15 gpurun_out/r02i_pytest.log
timeout 600 python bench.py --configs cfg2 --no-cpu --no-parity > gpurun_out/r02i_bench_cfg2.json 2> gpurun_out/r02i_bench.err; tail -3 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i_bench_cfg2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'u16', d.get('e2e_u16',{}).get('value'), 'pageable', d.get('e2e_pageable',{}).get('value'))
PY
for t in 4 8 16 24; do BLINGFIRE_B200_COPY_THREADS=$t timeout 600 python bench.py --configs cfg2 --no-cpu --no-parity --steps 8 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('copy threads $t pageable', d['e2e_pageable']['value'], 'e2e', d['e2e']['value'])"; done
