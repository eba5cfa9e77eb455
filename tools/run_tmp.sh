#!/bin/bash
mkdir -p gpurun_out/r2o
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_properties_gpu.py -x -q -m gpu -k "poisson or fuzz or propert or edge" 2>&1 | tail -5
for c in 2 3 4; do
  for m in dag phases; do
  echo "MCCNN_POISSON=$m"
  MCCNN_POISSON=$m timeout 300 python tools/conv_bench.py --config $c --op poisson_sampling --reps 20 2>&1 | tee gpurun_out/r2o/poisson_c${c}_$m.txt | tail -5
  done
done
for m in dag phases dag phases; do
MCCNN_POISSON=$m timeout 600 python bench.py --no-other-configs 2>gpurun_out/r2o/bench.err | tee gpurun_out/r2o/bench_$m.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$m', d['value'], d['ms_per_step'], d['latency']['ms_per_forward'], d['e2e']['value'], d['gpu_launches'])"
done
