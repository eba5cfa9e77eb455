#!/bin/bash
bash tools/gpu_multi.sh 8 r2final2_n8 bench
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
MCCNN_COLLECT=ingraph timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-other-configs 2>/dev/null > gpurun_out/r2final2_n8/bench_n8_s20_ingraph.json; cut -c1-260 gpurun_out/r2final2_n8/bench_n8_s20_ingraph.json
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-other-configs 2>/dev/null > gpurun_out/r2final2_n8/bench_n8_s20_b.json; cut -c1-260 gpurun_out/r2final2_n8/bench_n8_s20_b.json
