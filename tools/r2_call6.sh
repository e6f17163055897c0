#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c6; mkdir -p $O
timeout 500 python tools/batched_variants.py 0 3 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
cat $O/variants.log
timeout 600 python -m pytest tests/test_zz_gpu_refinement.py tests/test_gpu_kkt.py tests/test_gpu_batched.py -q > $O/pytest_part.log 2>&1; echo "pytest(part) rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_part.log
timeout 200 python tools/netlib_rate.py SCSD8 25FV47 TRUSS BANDM DEGEN2 GROW15 > $O/rate_pipe256.log 2>&1
IPM_PIPE_MIN_M=1024 timeout 200 python tools/netlib_rate.py SCSD8 25FV47 TRUSS BANDM DEGEN2 GROW15 > $O/rate_pipe1024.log 2>&1
paste -d'\n' $O/rate_pipe256.log $O/rate_pipe1024.log
IPM_PIPE_MIN_M=1024 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_netlib_all.py tests/test_gpu_boundary.py -q > $O/pytest_pipe1024.log 2>&1; echo "pytest(pipe1024) rc=$?" | tee -a $O/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_pipe1024.log | cut -c1-200
timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c6/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')}, d['e2e']['value'], d['config'].get('handed_to_augmented_system_kernel_rank0'))
print(d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
print(d['extras']['netlib']['QAP15'].get('dependent_rows_removed'))
PY
