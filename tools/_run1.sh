set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_r2_api.py -x -q -k "not sp" > gpurun_out/r02b_pytest.log 2>&1; tail -5 gpurun_out/r02b_pytest.log
timeout 600 python bench.py --configs cfg2 --no-cpu > gpurun_out/r02b_bench_cfg2.json 2> gpurun_out/r02b_bench.err; tail -3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench_cfg2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['match'], d['parity']['mismatching_docs'])
PY
bash tools/measure_round.sh r02b ncu2 > gpurun_out/r02b_measure.log 2>&1
