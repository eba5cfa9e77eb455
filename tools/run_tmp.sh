timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_model_gpu.py tests/test_models_gpu.py -x -q -k "neighbors or planned or model or Model" 2>&1 | tail -4
for c in 5 4 2; do timeout 300 python tools/conv_bench.py --config $c --op find_neighbors_fused --reps 10 2>&1 | tail -16 | cut -c1-140; done
