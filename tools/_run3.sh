set -x
timeout 1500 python -m pytest tests/test_gpu_parity_sp.py tests/test_gpu_offsets.py -x -q > gpurun_out/r02g_pytest.log 2>&1; tail -5 gpurun_out/r02g_pytest.log
timeout 900 python bench.py --configs cfg3 --no-cpu > gpurun_out/r02g_bench_cfg3.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench_cfg3.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['match'], d['parity']['mismatching_docs'], 'frac', d['roofline']['frac'])
PY
BLINGFIRE_B200_NO_MEMO=1 timeout 900 python bench.py --configs cfg3 --no-cpu --no-parity --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-memo value', d['value'])"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py gpt2.bin > gpurun_out/r02g_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02g_memcheck.log
bash tools/measure_round.sh r02g ncu3 > gpurun_out/r02g_measure.log 2>&1
