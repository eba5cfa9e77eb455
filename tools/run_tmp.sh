#!/bin/bash
mkdir -p gpurun_out/r2q
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_properties_gpu.py -x -q -m gpu -k "poisson or fuzz or propert or edge" 2>&1 | tail -5
for m in fit full fit full; do
MCCNN_POISSON_STAGE=$m timeout 600 python bench.py --no-other-configs 2>gpurun_out/r2q/bench.err | tee gpurun_out/r2q/bench_$m.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$m', d['value'], d['ms_per_step'], d['latency']['ms_per_forward'], d['e2e']['value'], d['gpu_launches'])"
done
for m in fit full; do
MCCNN_POISSON_STAGE=$m timeout 600 python bench.py --no-other-configs --steps 20 2>gpurun_out/r2q/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$m 20 steps', d['value'], d['ms_per_step'], d['latency']['ms_per_forward'], d['e2e']['value'], d['gpu_launches'])"
done
for c in 2 4 5; do
  for m in fit full; do
  echo "MCCNN_POISSON_STAGE=$m"
  MCCNN_POISSON_STAGE=$m timeout 300 python tools/conv_bench.py --config $c --op poisson_sampling --reps 20 2>&1 | tee gpurun_out/r2q/poisson_c${c}_$m.txt | tail -5
  done
done
