set -x
timeout 1500 python -m pytest tests/test_gpu_parity_sp.py tests/test_gpu_offsets.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r02l_pytest.log 2>&1; tail -6 gpurun_out/r02l_pytest.log
timeout 900 python bench.py --configs cfg4 --no-cpu > gpurun_out/r02l_bench_cfg4.json 2> gpurun_out/r02l_bench.err; tail -3 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_cfg4.json'))
print('cfg4 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['match'], d['parity']['mismatching_docs'], 'cold', d.get('cold_first_step'))
PY
BLINGFIRE_B200_NO_MEMO=1 timeout 900 python bench.py --configs cfg4 --no-cpu --no-parity --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-memo value', d['value'])"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py xlm_roberta_base.bin > gpurun_out/r02l_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02l_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py xlm_roberta_base.bin gpt2.bin > gpurun_out/r02l_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02l_racecheck.log
