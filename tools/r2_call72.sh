#!/bin/bash
# closing build of round 2: full GPU suite + bench N=1
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c72; mkdir -p $O
timeout 900 python -m pytest tests -q -x -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -2 $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity'] if 'parity' in d else '', d['roofline']['phase_ms_per_step'], d['roofline']['whole_step_frac'])"
