#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c42; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_parity.py tests/test_gpu_small_lp.py tests/test_zz_gpu_refinement.py -q -x > $O/pytest_part.log 2>&1; echo "pytest(part) rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_part.log
timeout 500 python tools/batched_variants.py 0 2 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
head -3 $O/variants.log | cut -c1-200
