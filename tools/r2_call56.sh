#!/bin/bash
# grids over the joined prefix: e2e trace, host-path tests, bench N=1
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c56; mkdir -p $O
timeout 300 python tools/e2e_trace.py > $O/e2e_trace.txt 2>&1; echo "trace rc=$?" | tee -a $O/summary.txt
cut -c1-300 $O/e2e_trace.txt | head -8
timeout 600 python -m pytest tests/test_gpu_batched.py -q -x > $O/pytest_batched.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_batched.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity'] if 'parity' in d else '')"
