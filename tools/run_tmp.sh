#!/bin/bash
mkdir -p gpurun_out/r2m
for rep in 1 2 3; do
for v in 0 8192 600; do
  echo "MCCNN_PDF_FUSE_MAX=$v"
  MCCNN_PDF_FUSE_MAX=$v timeout 600 python bench.py --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['latency']['ms_per_forward'], d['e2e']['value'])"
done
done
