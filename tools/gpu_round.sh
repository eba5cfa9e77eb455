#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms (default line incl. other_configs), sanitizer passes.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag] [quick]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q -s > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
cp gpurun_out/fullsize_parity_*.txt $out/ 2>/dev/null
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -2 $out/smoke.log
echo "== bench ours"
timeout 900 python bench.py --steps 200 --warmup 5 > $out/bench_ours.json 2> $out/bench_ours.err; tail -c 1500 $out/bench_ours.json; tail -5 $out/bench_ours.err
echo "== bench reference"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 1200 $out/bench_ref.json; tail -5 $out/bench_ref.err
if [ "$2" != "quick" ]; then
echo "== sanitizers over smoke()"
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > $out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> $out/sanitizer_$tool.log; tail -3 $out/sanitizer_$tool.log
done
echo "== sanitizers over the fuzz sweep and the newest kernels' parity tests (reference kernels excluded)"
SAN_TESTS="tests/test_fuzz_gpu.py::test_adversarial_geometry_bit_exact tests/test_parity_gpu.py::test_poisson_bit_exact_on_very_ragged_batches tests/test_parity_gpu.py::test_dense_fewrows_matches_torch tests/test_parity_gpu.py::test_find_neighbors_pdf_fused_equals_the_two_ops tests/test_parity_gpu.py::test_find_neighbors_fused_bit_exact"
for i in 0 1 2 3 4 5 6 7 8 9 10 11; do SAN_TESTS="$SAN_TESTS tests/test_fuzz_gpu.py::test_random_configuration[$i]"; done
for tool in memcheck racecheck; do
  MCCNN_NO_REF_KERNELS=1 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $SAN_TESTS -x -q -m gpu > $out/sanitizer_fuzz_$tool.log 2>&1
  echo "$tool exit $?" >> $out/sanitizer_fuzz_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/sanitizer_fuzz_$tool.log | tail -3
done
fi
ls -la $out
