timeout 300 python -m pytest tests/test_dense_tc_gpu.py -x -q -s 2>&1 | grep -v "^dense_tc " | tail -9
timeout 200 python tools/conv_bench.py --config 4 --op dense_layer_tc --reps 10 2>&1 | tail -11
