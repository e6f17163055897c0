#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c26; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_parity.py -q -x -k "syrk or potrf or batched or chol or dense" > $O/pytest_part.log 2>&1; echo "pytest(part) rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_part.log
timeout 500 python tools/batched_variants.py 0 3 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
head -4 $O/variants.log
timeout 300 python tools/syrk_probe.py > $O/syrk_probe.log 2>&1; grep syrk $O/syrk_probe.log
