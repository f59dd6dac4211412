#!/bin/bash
# end-to-end host pipeline of one configuration with the pipeline's own timing trace (stderr), for a few chunk sizes
cfg=${1:-cfg3}
for mb in 8 16 32 64; do
  echo "== chunk ${mb} MB"
  BLINGFIRE_B200_TRACE=1 BLINGFIRE_B200_CHUNK_MB=$mb python bench.py --configs $cfg --no-cpu --no-parity --steps 6 --warmup 3 2> /tmp/trace_$mb.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1f e2e %.2f GB/s (%.2f ms)'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
  grep "pipeline" /tmp/trace_$mb.err | tail -2
done
