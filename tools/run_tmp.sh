mkdir -p gpurun_out/r2j
TRB="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3
echo "== N=2 headline 20 steps"; timeout 600 $TRB bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j/bench_n2_s20.json 2> gpurun_out/r2j/bench_n2_s20.err; head -c 300 gpurun_out/r2j/bench_n2_s20.json; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" gpurun_out/r2j/bench_n2_s20.err | tail -5
echo "== N=2 headline"; timeout 600 $TRB bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2j/bench_n2.json 2> gpurun_out/r2j/bench_n2.err; head -c 300 gpurun_out/r2j/bench_n2.json
echo "== N=1 headline 20 steps"; timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs > gpurun_out/r2j/bench_n1_s20.json 2>/dev/null; head -c 300 gpurun_out/r2j/bench_n1_s20.json
