#!/bin/bash
# kb_chol with the factor-warp store: full GPU suite, bitwise fingerprint, bench N=1
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c55; mkdir -p $O
python tools/kbc_hash.py > $O/kbc_hash.txt 2>&1; tail -1 $O/kbc_hash.txt
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'] if 'parity' in d else '', d['roofline']['phase_ms_per_step'])"
