#!/bin/bash
# Round 2, second GPU call: the failing batched test under IPM_DEBUG_SYNC, the whole suite, the 65536-LP scan with the
# hand-off path, QAP probes, bench.
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c2; mkdir -p $O
IPM_DEBUG_SYNC=1 timeout 300 python -m pytest tests/test_gpu_batched.py -x -q -k "reference_goldens or single_lp_path" > $O/pytest_debug.log 2>&1; echo "debug rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest_debug.log
timeout 300 python -m pytest tests/test_gpu_kkt.py -q > $O/pytest_kkt.log 2>&1; echo "kkt rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_kkt.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -25 $O/pytest_gpu.log
timeout 400 python tools/scan_batch_gpu.py 0 1 2 3 4 5 6 7 --ab > $O/scan.log 2>&1; echo "scan rc=$?" | tee -a $O/summary.txt
tail -24 $O/scan.log
timeout 300 python tools/qap_probe.py QAP8 QAP12 QAP15 > $O/qap.log 2>&1; echo "qap rc=$?" | tee -a $O/summary.txt
cat $O/qap.log
timeout 500 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -c 600 $O/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/batched_launches.csv \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kb_chol -s 3 -c 1 -o $O/kb_chol \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kb_chol.log 2>&1
cat $O/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c2/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')}, d['e2e']['value'], d['config'].get('max_iteration_diff_vs_oracle_table'), d['config'].get('max_rel_objective_diff_vs_oracle_table'), d['config'].get('handed_to_augmented_system_kernel_rank0'))
print(d['roofline']['phase_ms_per_step'])
PY
