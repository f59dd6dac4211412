set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; tail -6 gpurun_out/r02j_pytest.log
python - <<'PY' > gpurun_out/r02j_load_times.txt 2>&1
import time, sys, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch; torch.zeros(1, device='cuda')
import blingfire_b200 as bf
from _common import model_path
for name in ("bert_base_tok.bin","bert_multi_cased.bin","gpt2.bin","xlm_roberta_base.bin"):
    t0=time.time(); h=bf.load_model(model_path(name)); torch.cuda.synchronize(); t1=time.time()
    print(name, "LoadModel %.2f s"%(t1-t0)); bf.free_model(h)
PY
cat gpurun_out/r02j_load_times.txt
timeout 600 python bench.py --configs cfg2 --no-cpu --no-parity > gpurun_out/r02j_bench_cfg2.json 2> gpurun_out/r02j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench_cfg2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'u16', d.get('e2e_u16',{}).get('value'), 'pageable', d.get('e2e_pageable',{}).get('value'))
PY
