#!/bin/bash
# One GPU-box pass that produces everything profiles/ holds for a round:
#   the -m gpu tests, the bench line (both arms), the other configurations, the ncu launch list of the
#   bench command and one --set full capture of the hot kernel.  Outputs go to gpurun_out/.
# Usage (from the repo root, on the GPU box):  bash tools/measure_round.sh [tag]
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
set -x
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
timeout 600 python tools/bench_configs.py --docs 500000 > $out/${tag}_other_configs.json 2> $out/${tag}_other_configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $out/${tag}_ncu_bench.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:wp_tokenize -c 1 -f -o $out/${tag}_wp_tokenize \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_pytest_gpu.log
cat $out/${tag}_bench.json
