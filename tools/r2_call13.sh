#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c14; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_parity.py -q -k "potrf or batched or chol or dense" > $O/pytest_part.log 2>&1; echo "pytest(part) rc=$?" | tee -a $O/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_part.log | cut -c1-200
timeout 500 python tools/batched_variants.py 0 2 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
head -3 $O/variants.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kb_chol -s 3 -c 1 -o $O/kb_chol \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kb_chol.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kbf_dir -s 4 -c 1 -o $O/kbf_dir \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kbf_dir.log 2>&1
