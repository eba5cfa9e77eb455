#!/bin/bash
# One gpurun call: GPU parity tests, both bench arms, ncu launch lists and one full capture of the hot kernels.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log
echo "== bench ours"
timeout 600 python bench.py --steps 200 --warmup 5 > $out/bench_ours.json 2> $out/bench_ours.err; tail -c 1500 $out/bench_ours.json
echo "== bench reference"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 600 $out/bench_ref.json
echo "== ncu launch list (eager planned forward)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/launches_forward.csv python tools/ncu_target.py --forwards 1 > $out/ncu_target.log 2>&1
python tools/launch_summary.py $out/launches_forward.csv > $out/launches_forward.summary.txt 2>&1; cat $out/launches_forward.summary.txt
echo "== ncu launch list (bench.py --steps 2 --warmup 1)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 1 > $out/ncu_bench.log 2>&1
python tools/launch_summary.py $out/launches_bench.csv > $out/launches_bench.summary.txt 2>&1; head -30 $out/launches_bench.summary.txt
echo "== ncu full capture of the hot kernels"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -c 30 -o $out/prof_hot -f python tools/ncu_target.py --forwards 1 > $out/ncu_full.log 2>&1
ls -la $out
