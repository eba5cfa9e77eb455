#!/bin/bash
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "persistent_conv" 2>&1 | tail -3
bash tools/gpu_multi.sh 2 r2final_n2 quick
