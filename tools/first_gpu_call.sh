#!/bin/bash
# One gpurun call that re-establishes the state of the tree on a B200 and brings back what the next kernel work needs:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
# Everything lands in gpurun_out/first_call/.  Nothing printed under ncu is a bench value.
set -u
export PYTHONPATH=$PWD
O=gpurun_out/first_call; mkdir -p $O
# 1. parity suite and smoke
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
# 2. bench, both arms
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 200 python bench.py --impl reference --steps 3 > $O/bench_reference.json 2> $O/bench_reference.err
# 3. stragglers of the weak-scaling workload beyond the three blocks examined in round 1 (seeds 24576..65535)
timeout 120 python tools/find_straggler.py 3 4 5 6 7 > $O/stragglers_restart_on.log 2>&1
timeout 200 python tools/find_straggler.py 3 4 5 6 7 --no-restart --alts > $O/stragglers_restart_off.log 2>&1
# 4. launch list of the bench command, then source-level captures of the two kernels with the most headroom
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dmma_ws_kernel -s 3 -c 1 -o $O/syrk_ws \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_syrk.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kb_chol -s 3 -c 1 -o $O/kb_chol \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kb_chol.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kbf_dir -s 4 -c 2 -o $O/kbf_dir \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kbf_dir.log 2>&1
ls -la $O | tee -a $O/summary.txt
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/stragglers_restart_on.log | tail -8
