set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_r2_api.py -x -q -k "not sp" > gpurun_out/r02e_pytest.log 2>&1; tail -5 gpurun_out/r02e_pytest.log
timeout 600 python bench.py --configs cfg2 --no-cpu > gpurun_out/r02e_bench_cfg2.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02e_bench_cfg2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'u16', d.get('e2e_u16',{}).get('value'), 'pageable', d.get('e2e_pageable',{}).get('value'), 'parity', d['parity']['match'], d['parity']['mismatching_docs'])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py bert_base_tok.bin > gpurun_out/r02e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02e_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py bert_base_tok.bin > gpurun_out/r02e_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02e_racecheck.log
