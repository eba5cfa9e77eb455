#!/bin/bash
mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_models_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "dense or model or mlp or planned or mcclass or fused" 2>&1 | tail -5
timeout 300 python tools/conv_bench.py --config 3 --op dense_layer_fewrows --reps 20 2>&1 | tee gpurun_out/r2r/fewrows_c3.txt | tail -9
timeout 600 python bench.py --config 3 --no-other-configs 2>gpurun_out/r2r/bench3.err | tee gpurun_out/r2r/bench3.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d.get('latency'), d['e2e']['value'], d['gpu_launches'])"
