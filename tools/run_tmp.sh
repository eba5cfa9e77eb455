mkdir -p gpurun_out/r2k
TRB="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
echo "== N=8 headline 200"; timeout 500 $TRB bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/r2k/bench_n8.json 2> gpurun_out/r2k/bench_n8.err; head -c 250 gpurun_out/r2k/bench_n8.json; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" gpurun_out/r2k/bench_n8.err | tail -5
echo "== N=8 headline 20"; timeout 500 $TRB bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2k/bench_n8_s20.json 2> gpurun_out/r2k/bench_n8_s20.err; head -c 250 gpurun_out/r2k/bench_n8_s20.json
echo "== N=8 config 4"; timeout 500 $TRB bench.py --gpus 8 --steps 20 --warmup 3 --config 4 > gpurun_out/r2k/bench_n8_c4.json 2> gpurun_out/r2k/bench_n8_c4.err; head -c 250 gpurun_out/r2k/bench_n8_c4.json; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" gpurun_out/r2k/bench_n8_c4.err | tail -5
echo "== N=8 config 5"; timeout 500 $TRB bench.py --gpus 8 --steps 20 --warmup 3 --config 5 > gpurun_out/r2k/bench_n8_c5.json 2> gpurun_out/r2k/bench_n8_c5.err; head -c 250 gpurun_out/r2k/bench_n8_c5.json; grep -v "OMP_NUM\|\*\*\*\*\|^$\|NCCL version" gpurun_out/r2k/bench_n8_c5.err | tail -5
