MCCNN_CONV_CFG=warp timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py -x -q 2>&1 | tail -2
for c in 2 5 4 3; do echo "== warp config $c"; MCCNN_CONV_CFG=warp timeout 200 python tools/conv_bench.py --config $c --reps 10 2>&1 | tail -19 | cut -c60-170; done
