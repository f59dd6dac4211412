timeout 900 python bench.py --configs cfg4 --no-cpu > gpurun_out/r02n_bench_cfg4.json 2> gpurun_out/r02n_bench.err; tail -3 gpurun_out/r02n_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_bench_cfg4.json'))
print('cfg4 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['match'], d['parity']['mismatching_docs'])
PY
bash tools/measure_round.sh r02n ncu4 > gpurun_out/r02n_measure.log 2>&1
