#!/bin/bash
# ncu --set full capture of the three Poisson launches of one forward + per-line stall summary
out=gpurun_out/${1:-poisson}
mkdir -p $out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "poisson or Poisson" 2>&1 | tail -2
timeout 300 python tools/quick_time.py 2>&1 | tail -25 > $out/quick_time.txt; cat $out/quick_time.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:poisson_kernel -c 3 \
    -o $out/prof -f python tools/ncu_target.py --forwards 1 > $out/ncu.log 2>&1
ncu -i $out/prof.ncu-rep --page source --csv --print-source cuda,sass > $out/src.csv 2>/dev/null
for i in 0 1; do python tools/ncu_source_hot.py $out/src.csv $i 28 > $out/hot$i.txt; cat $out/hot$i.txt; done
rm -f $out/src.csv
