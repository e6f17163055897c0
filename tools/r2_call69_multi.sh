#!/bin/bash
# Round 2: 8-GPU run of the bench command exactly as the driver launches it (weak scaling line + config.strong), then the
# reference arm under torchrun.   gpurun --gpus 8 --timeout 1200 -- 'bash tools/r2_call5_multi.sh'
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c69; mkdir -p $O
N=${1:-8}
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?" | tee -a $O/summary.txt
tail -c 1200 $O/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('$O/bench_n$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','parity','n_gpus','per_rank_ms_per_step')})
print('e2e', d['e2e'])
print('config', {k:v for k,v in d['config'].items() if k not in ('workload',)})
PY
