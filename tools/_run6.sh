set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sentences_batch or words_batch or cfg1" > gpurun_out/r02k_pytest.log 2>&1; tail -6 gpurun_out/r02k_pytest.log
timeout 900 python bench.py --no-cpu > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -3 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench.json'))
def show(n,r): print(n,'value %.1f'%r['value'],'e2e %.1f'%r['e2e']['value'],'cold', r.get('cold_first_step'), 'parity', r['parity']['match'])
show('cfg2',d)
for k,v in d['other_configs'].items(): show(k,v)
PY
