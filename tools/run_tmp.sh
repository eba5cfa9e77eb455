#!/bin/bash
# Scratch command file for one gpurun call (overwritten per experiment):  gpurun -- 'bash tools/run_tmp.sh'
# The reusable sequences are tools/gpu_round.sh (single GPU) and tools/gpu_multi.sh (N GPUs).
bash tools/gpu_round.sh scratch quick
