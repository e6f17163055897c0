#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c19; mkdir -p $O
./tools/dmma_operand_probe | tee $O/dmma_operand_probe.txt
timeout 900 python -m pytest tests/test_gpu_batched.py -q -x -k "rhs_from_syrk or syrk" > $O/pytest_part.log 2>&1; echo "pytest(part) rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_part.log
timeout 500 python tools/batched_variants.py 0 2 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
head -3 $O/variants.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dmma_ws -s 3 -c 1 -o $O/syrk \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_syrk.log 2>&1; echo "ncu syrk rc=$?" | tee -a $O/summary.txt
