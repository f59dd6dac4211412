#!/bin/bash
# One GPU-box pass that produces everything profiles/ holds for a round:
#   the -m gpu tests, the bench line (both arms), the ncu launch list of the bench command and one
#   --set full capture (with source) of each hot kernel.  Outputs go to gpurun_out/.
# Usage (from the repo root, on the GPU box):  bash tools/measure_round.sh [tag] [what...]
#   what: tests bench ref launches ncu2 ncu3 ncu4   (default: all)
tag=${1:-r02}
shift
what=${@:-tests bench ref launches ncu2 ncu3 ncu4}
out=gpurun_out
mkdir -p $out
has() { [[ " $what " == *" $1 "* ]]; }
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_box.txt 2>&1
lscpu | egrep 'Model name|^CPU\(s\)|NUMA' >> $out/${tag}_box.txt
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
  tail -5 $out/${tag}_pytest_gpu.log
fi
if has bench; then
  timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
  tail -3 $out/${tag}_bench.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
fi
if has launches; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --configs cfg2 > $out/${tag}_ncu_bench.log 2>&1
fi
ncu_full() {  # cfg kernel-regex name
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$2 -s 2 -c 1 -f -o $out/${tag}_$3 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-parity --configs $1 > $out/${tag}_ncu_$3.log 2>&1
  ncu -i $out/${tag}_$3.ncu-rep --page raw --csv > $out/${tag}_$3_raw.csv 2>/dev/null
  ncu -i $out/${tag}_$3.ncu-rep --page source --csv > $out/${tag}_$3_source.csv 2>/dev/null
}
has ncu2 && ncu_full cfg2 wp_tokenize wp_tokenize
has ncu3 && ncu_full cfg3 sp_bpe sp_bpe
has ncu4 && ncu_full cfg4 sp_unigram sp_unigram
has bench && cat $out/${tag}_bench.json
ls -la $out
