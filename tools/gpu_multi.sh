#!/bin/bash
# One gpurun --gpus N call: multi-rank correctness checks, then the bench lines of the N-GPU evidence.
# Usage (repo root, on the GPU box):  bash tools/gpu_multi.sh N tag [quick|bench]   (quick: no configs 4/5; bench: only
# the peer check and the weak-scaling lines)
N=${1:-2}
tag=${2:-multi}
out=gpurun_out/$tag
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
echo "== peer exchange check"
timeout 300 $TR tools/peer_gather_check.py > $out/peer_check.log 2>&1; echo "exit $?" >> $out/peer_check.log; tail -3 $out/peer_check.log
if [ "$3" != "bench" ]; then
echo "== sharded segmentation check"
timeout 600 $TR tools/sharded_seg_check.py > $out/sharded_seg.log 2>&1; echo "exit $?" >> $out/sharded_seg.log; tail -3 $out/sharded_seg.log
fi
echo "== bench weak, 200 steps"
timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 5 --no-other-configs > $out/bench_n$N.json 2> $out/bench_n$N.err; cut -c1-260 $out/bench_n$N.json
echo "== bench weak, 20 steps (the driver's command)"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-other-configs > $out/bench_n${N}_s20.json 2> $out/bench_n${N}_s20.err; cut -c1-260 $out/bench_n${N}_s20.json
if [ "$N" -le 2 ]; then   # the single-GPU reference points of the same box (skipped where N x box time is charged)
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs > $out/bench_n1_s20.json 2> $out/bench_n1_s20.err; cut -c1-260 $out/bench_n1_s20.json
timeout 600 python bench.py --steps 200 --warmup 5 --no-other-configs > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-260 $out/bench_n1.json
fi
if [ "$3" != "quick" ] && [ "$3" != "bench" ]; then
for c in 4 5; do
  echo "== config $c on $N GPUs"
  timeout 600 $TR bench.py --gpus $N --config $c --steps 20 --warmup 3 > $out/bench_n${N}_c$c.json 2> $out/bench_n${N}_c$c.err; cut -c1-260 $out/bench_n${N}_c$c.json; tail -2 $out/bench_n${N}_c$c.err
done
fi
ls $out
