#!/bin/bash
# end-to-end path against the H2D chunk size (LPs per chunk), closing build
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c73; mkdir -p $O
for c in 64 128 256; do
  IPM_E2E_CHUNK=$c timeout 120 python tools/e2e_trace.py > $O/e2e_trace_chunk$c.txt 2>&1
  echo "chunk $c: $(grep -h 'end to end' $O/e2e_trace_chunk$c.txt | cut -c1-120)"
done
