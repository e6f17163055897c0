#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c16; mkdir -p $O
timeout 300 python tools/batched_variants.py 0 3 1024 > $O/variants_b1024.log 2>&1; echo "variants1024 rc=$?" | tee -a $O/summary.txt
cat $O/variants_b1024.log | tail -8
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dmma_ws -s 3 -c 1 -o $O/syrk \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_syrk.log 2>&1; echo "ncu syrk rc=$?" | tee -a $O/summary.txt
